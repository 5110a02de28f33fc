// Host versions of the reference's quadratic-form helpers (ML/LinearAlgebra.cpp:8-73): same
// contracts and error behaviour; plain loops over the upper triangle.
#include "ML/LinearAlgebra.hpp"

#include <stdexcept>

namespace ml
{
	namespace LinearAlgebra
	{
		double xAx_symmetric(const Eigen::MatrixXd& A, PointView x)
		{
			const Eigen::Index n = A.rows();
			if (A.cols() != n) {
				throw std::invalid_argument("A is not square");
			}
			if (x.size() != n) {
				throw std::invalid_argument("Size mismatch");
			}
			const double* a = A.data();
			const double* v = x.data();
			double total = 0;
			for (Eigen::Index col = 0; col < n; ++col) {
				const double v_col = v[col];
				const double* column = a + col * n;
				total += column[col] * v_col * v_col;
				for (Eigen::Index row = 0; row < col; ++row) {
					total += 2 * column[row] * v_col * v[row];
				}
			}
			return total;
		}

		void xxT(PointView x, Eigen::MatrixXd& dest)
		{
			const Eigen::Index n = x.size();
			if (dest.rows() != n || dest.cols() != n) {
				dest.resize(n, n);
			}
			const double* v = x.data();
			double* out = dest.data();
			for (Eigen::Index col = 0; col < n; ++col) {
				out[col + col * n] = v[col] * v[col];
				for (Eigen::Index row = 0; row < col; ++row) {
					const double product = v[col] * v[row];
					out[row + col * n] = product;
					out[col + row * n] = product;
				}
			}
		}

		void add_a_xxT(PointView x, Eigen::MatrixXd& dest, const double a)
		{
			const Eigen::Index n = dest.rows();
			if (dest.cols() != n) {
				throw std::invalid_argument("Matrix is not square");
			}
			if (x.size() != n) {
				throw std::invalid_argument("Size mismatch");
			}
			const double* v = x.data();
			double* out = dest.data();
			for (Eigen::Index col = 0; col < n; ++col) {
				out[col + col * n] += a * v[col] * v[col];
				for (Eigen::Index row = 0; row < col; ++row) {
					const double term = a * v[col] * v[row];
					out[row + col * n] += term;
					out[col + row * n] += term;
				}
			}
		}
	}
}
