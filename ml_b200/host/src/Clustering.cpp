// Host-side initialisers with the reference's semantics (ML/Clustering.cpp:16-89).  They draw from
// the caller's std::default_random_engine with the same std:: facilities in the same order, so a
// seeded model starts from exactly the reference's initial state.
#include "ML/Clustering.hpp"

#include <algorithm>
#include <iterator>
#include <limits>
#include <numeric>
#include <stdexcept>

namespace ml
{
	namespace Clustering
	{
		namespace
		{
			inline const double* point(DataView data, Eigen::Index i)
			{
				return data.data() + i * data.outerStride();
			}

			inline double* column(MatrixOut m, Eigen::Index j)
			{
				return m.data() + j * m.outerStride();
			}

			/** 0, 1, 2, ... as an iterator (random access only so that std::distance is O(1); std::sample walks it forwards). */
			struct CountingIterator
			{
				using iterator_category = std::random_access_iterator_tag;
				using value_type = Eigen::Index;
				using difference_type = std::ptrdiff_t;
				using pointer = const Eigen::Index*;
				using reference = const Eigen::Index&;
				Eigen::Index value;
				reference operator*() const { return value; }
				CountingIterator& operator++() { ++value; return *this; }
				CountingIterator operator++(int) { CountingIterator old = *this; ++value; return old; }
				bool operator==(const CountingIterator& other) const { return value == other.value; }
				bool operator!=(const CountingIterator& other) const { return value != other.value; }
				difference_type operator-(const CountingIterator& other) const { return value - other.value; }
			};

			inline double squared_distance(const double* a, const double* b, Eigen::Index dim)
			{
				double total = 0;
				for (Eigen::Index l = 0; l < dim; ++l) {
					const double diff = a[l] - b[l];
					total += diff * diff;
				}
				return total;
			}
		}

		Model::~Model() = default;

		CentroidsInitialiser::~CentroidsInitialiser() = default;

		ResponsibilitiesInitialiser::~ResponsibilitiesInitialiser() = default;

		void Forgy::init(DataView data, Prng& prng, const unsigned int number_components, MatrixOut centroids) const
		{
			// std::sample over 0..N-1 (selection sampling): the draw sequence the reference makes.  The population is
			// a counting iterator instead of the reference's materialised index vector (800 MB at N = 1e8): std::sample
			// only walks it forwards, so the draws and the chosen indices are the same.
			std::vector<Eigen::Index> chosen;
			chosen.reserve(number_components);
			std::sample(CountingIterator{0}, CountingIterator{data.cols()}, std::back_inserter(chosen), number_components, prng);
			for (unsigned int k = 0; k < number_components; ++k) {
				std::copy_n(point(data, chosen[k]), data.rows(), column(centroids, k));
			}
		}

		void RandomPartition::init(DataView data, Prng& prng, const unsigned int number_components, MatrixOut centroids) const
		{
			const Eigen::Index dim = data.rows();
			for (unsigned int k = 0; k < number_components; ++k) {
				std::fill_n(column(centroids, k), dim, 0.0);
			}
			std::vector<unsigned int> members(number_components, 0u);
			std::uniform_int_distribution<unsigned int> pick(0, number_components - 1);
			for (Eigen::Index i = 0; i < data.cols(); ++i) {
				const unsigned int k = pick(prng);
				const double count = static_cast<double>(++members[k]);
				double* c = column(centroids, k);
				const double* x = point(data, i);
				for (Eigen::Index l = 0; l < dim; ++l) {
					c[l] += (x[l] - c[l]) / count;   // running mean of the group
				}
			}
		}

		void KPP::init(DataView data, Prng& prng, const unsigned int number_components, MatrixOut centroids) const
		{
			// Weight of a point = squared distance to the nearest centroid chosen so far (1 for the
			// first draw).  The minimum is kept incrementally: one pass over the data per new centroid,
			// O(N K D) in total, with the same values the reference recomputes from scratch.
			const Eigen::Index dim = data.rows();
			const Eigen::Index n = data.cols();
			std::vector<double> nearest(static_cast<size_t>(n), std::numeric_limits<double>::infinity());
			std::vector<double> weights(static_cast<size_t>(n), 1.0);
			for (unsigned int k = 0; k < number_components; ++k) {
				if (k > 0) {
					const double* newest = column(centroids, k - 1);
					for (Eigen::Index i = 0; i < n; ++i) {
						nearest[static_cast<size_t>(i)] = std::min(nearest[static_cast<size_t>(i)], squared_distance(point(data, i), newest, dim));
						weights[static_cast<size_t>(i)] = nearest[static_cast<size_t>(i)];
					}
				}
				std::discrete_distribution<Eigen::Index> draw(weights.begin(), weights.end());
				const Eigen::Index index = draw(prng);
				std::copy_n(point(data, index), dim, column(centroids, k));
			}
		}

		ExplicitCentroids::ExplicitCentroids(Eigen::MatrixXd centroids)
			: centroids_(std::move(centroids))
		{
		}

		void ExplicitCentroids::init(DataView data, Prng&, const unsigned int number_components, MatrixOut centroids) const
		{
			if (centroids_.rows() != data.rows() || centroids_.cols() != static_cast<Eigen::Index>(number_components)) {
				throw std::invalid_argument("ExplicitCentroids: stored centroids have the wrong shape");
			}
			for (unsigned int k = 0; k < number_components; ++k) {
				std::copy_n(centroids_.data() + static_cast<Eigen::Index>(k) * centroids_.rows(), centroids_.rows(), column(centroids, k));
			}
		}

		ClosestCentroid::ClosestCentroid(std::shared_ptr<const CentroidsInitialiser> centroids_initialiser)
			: centroids_initialiser_(centroids_initialiser)
		{
			if (!centroids_initialiser_) {
				throw std::invalid_argument("Null centroids initialiser");
			}
		}

		void ClosestCentroid::init(DataView data, Prng& prng, unsigned int number_components, MatrixOut responsibilities) const
		{
			const Eigen::Index dim = data.rows();
			Eigen::MatrixXd centroids(dim, number_components);
			centroids_initialiser_->init(data, prng, number_components, centroids);
			for (unsigned int k = 0; k < number_components; ++k) {
				std::fill_n(column(responsibilities, k), data.cols(), 0.0);
			}
			for (Eigen::Index i = 0; i < data.cols(); ++i) {
				const double* x = point(data, i);
				unsigned int winner = 0;
				double smallest = squared_distance(x, centroids.data(), dim);
				for (unsigned int k = 1; k < number_components; ++k) {
					const double candidate = squared_distance(x, centroids.data() + static_cast<Eigen::Index>(k) * dim, dim);
					if (candidate < smallest) {   // strict: the lowest index wins ties
						smallest = candidate;
						winner = k;
					}
				}
				column(responsibilities, winner)[i] = 1.0;
			}
		}
	}
}
