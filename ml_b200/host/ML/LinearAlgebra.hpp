#pragma once
// Host-side helpers with the reference's names and contracts (ML/LinearAlgebra.hpp:12-31).  On the
// device they are absorbed into the EM kernels; these versions serve EM::assign_responsibilities
// and callers of the public API.
#include <Eigen/Core>
#include "Clustering.hpp"   // PointView
#include "dll.hpp"

namespace ml
{
	namespace LinearAlgebra
	{
		/** x^T A x for symmetric A, reading only the upper triangle of A.
		@throw std::invalid_argument If A is not square or x has the wrong size. */
		DLL_DECLSPEC double xAx_symmetric(const Eigen::MatrixXd& A, PointView x);

		/** dest = x x^T (dest is resized if needed). */
		DLL_DECLSPEC void xxT(PointView x, Eigen::MatrixXd& dest);

		/** dest += a x x^T.
		@throw std::invalid_argument If dest is not square or x has the wrong size. */
		DLL_DECLSPEC void add_a_xxT(PointView x, Eigen::MatrixXd& dest, double a);
	}
}
