#pragma once
// Private glue between the ML++ host classes and the C-ABI of include/mlb200.h: RAII handles, the
// process-wide GPU context, and the mapping of status codes to the exceptions the reference throws.
#include <cstdint>
#include <memory>
#include <vector>
#include <Eigen/Core>

extern "C" {
#include "mlb200.h"
}

namespace ml
{
	namespace detail
	{
		/** Throws std::invalid_argument for MLB_EINVAL and std::runtime_error for every other failure. */
		void check(int status, const char* where);

		/** The process-wide context: MLPP_CUDA_DEVICES = number of GPUs to shard over (1, 2, 4 or 8;
		default 1).  Created on first use; there is no CPU fallback, so a machine without a CUDA device
		gets a std::runtime_error from the first fit(). */
		mlb_ctx* shared_context();

		/** The point matrix in HBM. */
		class DeviceData
		{
		public:
			explicit DeviceData(Eigen::Ref<const Eigen::MatrixXd> data);
			~DeviceData();
			DeviceData(const DeviceData&) = delete;
			DeviceData& operator=(const DeviceData&) = delete;
			mlb_data* handle() const { return handle_; }
			Eigen::Index rows() const { return rows_; }
			Eigen::Index cols() const { return cols_; }
		private:
			mlb_data* handle_ = nullptr;
			Eigen::Index rows_ = 0, cols_ = 0;
		};

		/** Device state of one EM fit. */
		class EmDevice
		{
		public:
			EmDevice(Eigen::Ref<const Eigen::MatrixXd> data, unsigned int number_components);
			~EmDevice();
			EmDevice(const EmDevice&) = delete;
			EmDevice& operator=(const EmDevice&) = delete;
			Eigen::MatrixXd sample_covariance();
			void set_parameters(const Eigen::MatrixXd& means, const std::vector<Eigen::MatrixXd>& covariances, const Eigen::VectorXd& mixing_probabilities);
			void maximise_from(const Eigen::MatrixXd& responsibilities);
			double step();
			void get_parameters(Eigen::MatrixXd& means, std::vector<Eigen::MatrixXd>& covariances, Eigen::VectorXd& mixing_probabilities);
			void emit(Eigen::MatrixXd* responsibilities, std::vector<unsigned int>* labels);
			void emit_rows(Eigen::Index begin, Eigen::Index count, Eigen::MatrixXd& responsibilities);
		private:
			DeviceData data_;
			mlb_em* em_ = nullptr;
			unsigned int number_components_;
		};

		/** Device state of one K-means fit (shared by all initialisations of a multi-start fit). */
		class KmDevice
		{
		public:
			KmDevice(Eigen::Ref<const Eigen::MatrixXd> data, unsigned int number_clusters);
			~KmDevice();
			KmDevice(const KmDevice&) = delete;
			KmDevice& operator=(const KmDevice&) = delete;
			void set_centroids(const Eigen::MatrixXd& centroids);
			void get_centroids(Eigen::MatrixXd& centroids);
			/** @return inertia; `changed` receives the number of labels that differ from the previous assignment. */
			double assign(std::int64_t& changed);
			/** @return squared Frobenius norm of the centroid shift. */
			double update();
			void get_labels(std::vector<unsigned int>& labels);
		private:
			DeviceData data_;
			mlb_km* km_ = nullptr;
		};
	}
}
