#pragma once
// ml::Clustering interfaces and initialisers: the same public surface as the reference's
// ML/Clustering.hpp:17-126.  Initialisers run on the host, once per fit, with the caller's matrix
// and the model's std::default_random_engine, so a fit started from them sees exactly the
// reference's pseudo-random stream.
#include <memory>
#include <random>
#include <vector>
#include <Eigen/Core>
#include "dll.hpp"

namespace ml
{
	/** The matrix views every entry point of this API takes (aliases only: the types, and so the signatures, are the
	reference's `Eigen::Ref<const Eigen::MatrixXd>`, `Eigen::Ref<Eigen::MatrixXd>` and their vector counterparts). */
	using DataView = Eigen::Ref<const Eigen::MatrixXd>;
	using MatrixOut = Eigen::Ref<Eigen::MatrixXd>;
	using PointView = Eigen::Ref<const Eigen::VectorXd>;
	using VectorOut = Eigen::Ref<Eigen::VectorXd>;
	using Prng = std::default_random_engine;

	namespace Clustering
	{
		/** A clustering model: fit() on a column-major matrix with one data point per column. */
		class Model
		{
		public:
			DLL_DECLSPEC virtual ~Model();

			/** Fits the model.
			@param data D x N matrix, a point in every column.
			@return Whether the fit converged.
			@throw std::invalid_argument If `data` has no rows or fewer columns than clusters.
			*/
			virtual bool fit(DataView data) = 0;

			virtual unsigned int number_clusters() const = 0;

			/** Labels of the fitted points (valid after a converged fit). */
			virtual const std::vector<unsigned int>& labels() const = 0;

			/** D x number_clusters() matrix of cluster centres. */
			virtual const Eigen::MatrixXd& centroids() const = 0;

			virtual bool converged() const = 0;
		};

		/** Chooses initial centroid locations. */
		class CentroidsInitialiser
		{
		public:
			DLL_DECLSPEC virtual ~CentroidsInitialiser();

			/**
			@param[in] data D x N data.
			@param[in,out] prng The model's generator.
			@param[in] number_components K <= N.
			@param[out] centroids D x K destination.
			*/
			DLL_DECLSPEC virtual void init(DataView data, Prng& prng, unsigned int number_components, MatrixOut centroids) const = 0;
		};

		/** Chooses initial responsibilities (N x K). */
		class ResponsibilitiesInitialiser
		{
		public:
			DLL_DECLSPEC virtual ~ResponsibilitiesInitialiser();

			DLL_DECLSPEC virtual void init(DataView data, Prng& prng, unsigned int number_components, MatrixOut responsibilities) const = 0;
		};

		/** K distinct data points drawn without replacement. */
		class Forgy : public CentroidsInitialiser
		{
		public:
			DLL_DECLSPEC void init(DataView data, Prng& prng, unsigned int number_components, MatrixOut centroids) const override;
		};

		/** Means of a uniformly random partition of the points into K groups. */
		class RandomPartition : public CentroidsInitialiser
		{
		public:
			DLL_DECLSPEC void init(DataView data, Prng& prng, unsigned int number_components, MatrixOut centroids) const override;
		};

		/** K-means++ seeding. */
		class KPP : public CentroidsInitialiser
		{
		public:
			DLL_DECLSPEC void init(DataView data, Prng& prng, unsigned int number_components, MatrixOut centroids) const override;
		};

		/** The given D x K matrix, verbatim (an addition to the reference's set): lets a caller start every rank of a
		multi-GPU job, or a benchmark, from exactly the same centroids without touching the pseudo-random stream. */
		class ExplicitCentroids : public CentroidsInitialiser
		{
		public:
			DLL_DECLSPEC explicit ExplicitCentroids(Eigen::MatrixXd centroids);

			/** @throw std::invalid_argument If the stored matrix is not D x number_components. */
			DLL_DECLSPEC void init(DataView data, Prng& prng, unsigned int number_components, MatrixOut centroids) const override;
		private:
			Eigen::MatrixXd centroids_;
		};

		/** One-hot responsibilities: every point belongs to its nearest initial centroid. */
		class ClosestCentroid : public ResponsibilitiesInitialiser
		{
		public:
			/** @throw std::invalid_argument If `centroids_initialiser` is null. */
			DLL_DECLSPEC ClosestCentroid(std::shared_ptr<const CentroidsInitialiser> centroids_initialiser);

			DLL_DECLSPEC void init(DataView data, Prng& prng, unsigned int number_components, MatrixOut responsibilities) const override;

			/** The initialiser the centroids come from (an addition: lets ml::EM run the nearest-centroid pass on the device). */
			const std::shared_ptr<const CentroidsInitialiser>& centroids_initialiser() const
			{
				return centroids_initialiser_;
			}
		private:
			std::shared_ptr<const CentroidsInitialiser> centroids_initialiser_;
		};
	}
}
