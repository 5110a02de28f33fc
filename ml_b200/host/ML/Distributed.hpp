#pragma once
// One process per GPU (torchrun / MPI style) for the ML++ clustering classes on the B200 backend.  An addition:
// the reference is a single-threaded CPU library and has no counterpart.
//
// Every rank calls ml::Distributed::init once, then uses ml::EM / ml::Clustering::KMeans as usual, passing ITS OWN
// columns of the data to fit(): rank r of `world` holds the global point range shard_range(n_total, world, r).
// Parameters (means, covariances, weights, centroids, log-likelihood, inertia, iteration count) are identical on every
// rank and bitwise equal to a single-GPU fit of the whole matrix; labels() and responsibilities() cover the rank's own
// points.  The initial state must be the same on every rank: use an initialiser that does not look at the local
// columns (Clustering::ExplicitCentroids), since the built-in ones draw from the matrix they are given.
#include <array>
#include <utility>
#include <Eigen/Core>
#include "dll.hpp"

namespace ml
{
	namespace Distributed
	{
		/** The 128-byte NCCL id rank 0 creates and ships to the other ranks (any side channel). */
		DLL_DECLSPEC std::array<unsigned char, 128> unique_id();

		/** Replaces the process-wide GPU context by rank `rank` of `world` (1, 2, 4 or 8) on CUDA device `device`.
		@throw std::logic_error If a context already exists (call before the first fit).
		@throw std::invalid_argument / std::runtime_error As the C-ABI reports. */
		DLL_DECLSPEC void init(int device, int rank, int world, const std::array<unsigned char, 128>& id);

		/** Whether init() has been called. */
		DLL_DECLSPEC bool active();

		/** Half-open global point range [begin, end) that `rank` of `world` holds for `n_total` points. */
		DLL_DECLSPEC std::pair<Eigen::Index, Eigen::Index> shard_range(Eigen::Index n_total, int world, int rank);
	}
}
