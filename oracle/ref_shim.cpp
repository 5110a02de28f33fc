// TEST INFRASTRUCTURE ONLY — C entry points (ctypes) onto the reference's OWN classes, compiled from the sources
// under /root/reference/ML by oracle/Makefile into oracle/_ref/libmlpp_ref.so (never copied into this repository).
// The only non-reference code in that library is this file and the Eigen stand-in under oracle/eigen_standin/.
//
// The option/result structs are the ones of mlpp_oracle.cpp so that oracle/__init__.py drives both libraries the same
// way.  What the reference does not expose is reported as follows:
//   - the iteration count of ml::EM::fit: counted from the "Step <n>" lines its verbose mode prints (EM.cpp:149-150),
//     with std::cout redirected into a counting buffer for the duration of the fit;
//   - the iteration count of KMeans::fit: the smallest maximum_steps for which the fit still converges (bisection;
//     every run re-seeds the PRNG, so runs are identical up to the step limit);
//   - inverse covariances / sqrt determinants are private to ml::EM and are not returned.
#include <chrono>
#include <cstdint>
#include <cstring>
#include <iostream>
#include <memory>
#include <streambuf>
#include <string>

#include "Clustering.hpp"
#include "EM.hpp"
#include "KMeans.hpp"
#include "LinearAlgebra.hpp"

namespace {

class ExplicitMeans : public ml::Clustering::CentroidsInitialiser {
public:
    ExplicitMeans(const double* means, Eigen::Index d, unsigned k) : means_(means), d_(d), k_(k) {}
    void init(Eigen::Ref<const Eigen::MatrixXd>, std::default_random_engine&, unsigned int number_components,
              Eigen::Ref<Eigen::MatrixXd> centroids) const override
    {
        for (unsigned c = 0; c < number_components && c < k_; ++c)
            for (Eigen::Index i = 0; i < d_; ++i) centroids(i, c) = means_[i + c * d_];
    }

private:
    const double* means_;
    Eigen::Index d_;
    unsigned k_;
};

std::shared_ptr<const ml::Clustering::CentroidsInitialiser> make_initialiser(int kind, const double* explicit_means, Eigen::Index d, unsigned k)
{
    switch (kind) {
    case 0: return std::make_shared<ml::Clustering::Forgy>();
    case 1: return std::make_shared<ml::Clustering::RandomPartition>();
    case 2: return std::make_shared<ml::Clustering::KPP>();
    default: return std::make_shared<ExplicitMeans>(explicit_means, d, k);
    }
}

struct StepCounter : std::streambuf {
    size_t steps = 0;
    std::string line;
    int overflow(int c) override
    {
        if (c == '\n') {
            if (line.rfind("Step ", 0) == 0) ++steps;
            line.clear();
        } else if (line.size() < 8) {
            line.push_back(static_cast<char>(c));
        }
        return c;
    }
};

Eigen::Ref<const Eigen::MatrixXd> wrap(const double* data, int64_t d, int64_t n, int64_t ld)
{
    return Eigen::Ref<const Eigen::MatrixXd>(Eigen::ConstView(data, d, n, 1, ld));
}

double seconds_since(std::chrono::steady_clock::time_point t0)
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

}  // namespace

extern "C" {

struct mlpp_oracle_em_options {
    unsigned seed;
    int set_seed;
    double absolute_tolerance;
    double relative_tolerance;
    unsigned maximum_steps;
    int means_init_kind;
    int resp_init_centroid_kind;
    int maximise_first;
    const double* explicit_means;
};

struct mlpp_oracle_em_result {
    double* means;
    double* covariances;
    double* mixing_probabilities;
    double* responsibilities;
    unsigned* labels;
    double* inverse_covariances;            // not available from the reference: left untouched
    double* sqrt_covariance_determinants;   // not available from the reference: left untouched
    double* step_seconds;                   // entry 0 receives the whole-fit seconds (the reference has no per-step clock)
    double log_likelihood;
    int converged;
    unsigned iterations;
};

// ml::EM::fit.  count_iterations != 0 runs the fit in verbose mode with std::cout captured (see the header comment).
// queries (d x n_queries, may be NULL) are passed to EM::assign_responsibilities after the fit; query_out is k x n_queries.
int mlpp_ref_em_fit(const double* data, int64_t d, int64_t n, int64_t ld, unsigned k, const mlpp_oracle_em_options* opt,
                    mlpp_oracle_em_result* out, int count_iterations, const double* queries, int64_t n_queries, double* query_out)
{
    try {
        ml::EM em(k);
        if (opt->set_seed) em.set_seed(opt->seed);
        em.set_absolute_tolerance(opt->absolute_tolerance);
        em.set_relative_tolerance(opt->relative_tolerance);
        em.set_maximum_steps(opt->maximum_steps);
        em.set_means_initialiser(make_initialiser(opt->means_init_kind, opt->explicit_means, d, k));
        em.set_responsibilities_initialiser(std::make_shared<ml::Clustering::ClosestCentroid>(
            make_initialiser(opt->resp_init_centroid_kind, opt->explicit_means, d, k)));
        em.set_maximise_first(opt->maximise_first != 0);
        StepCounter counter;
        std::streambuf* saved = nullptr;
        if (count_iterations) {
            em.set_verbose(true);
            saved = std::cout.rdbuf(&counter);
        }
        const auto t0 = std::chrono::steady_clock::now();
        bool converged = false;
        try {
            converged = em.fit(wrap(data, d, n, ld));
        } catch (...) {
            if (saved) std::cout.rdbuf(saved);
            throw;
        }
        const double secs = seconds_since(t0);
        if (saved) std::cout.rdbuf(saved);
        const size_t D = static_cast<size_t>(d), K = k, N = static_cast<size_t>(n);
        if (out->means) std::memcpy(out->means, em.means().data(), sizeof(double) * D * K);
        for (size_t c = 0; c < K; ++c)
            if (out->covariances && static_cast<size_t>(em.covariances()[c].size()) == D * D)
                std::memcpy(out->covariances + c * D * D, em.covariances()[c].data(), sizeof(double) * D * D);
        if (out->mixing_probabilities) std::memcpy(out->mixing_probabilities, em.mixing_probabilities().data(), sizeof(double) * K);
        if (out->responsibilities) std::memcpy(out->responsibilities, em.responsibilities().data(), sizeof(double) * N * K);
        if (out->labels) std::memcpy(out->labels, em.labels().data(), sizeof(unsigned) * N);
        if (out->step_seconds) out->step_seconds[0] = secs;
        out->log_likelihood = em.log_likelihood();
        out->converged = converged ? 1 : 0;
        out->iterations = count_iterations ? static_cast<unsigned>(counter.steps) : 0;
        for (int64_t q = 0; queries && q < n_queries; ++q) {
            Eigen::VectorXd x(d), u(k);
            for (int64_t i = 0; i < d; ++i) x[i] = queries[i + q * d];
            em.assign_responsibilities(x, u);
            for (unsigned c = 0; c < k; ++c) query_out[c + q * k] = u[c];
        }
        return 0;
    } catch (const std::exception&) {
        return 1;
    }
}

struct mlpp_oracle_kmeans_options {
    unsigned seed;
    int set_seed;
    double absolute_tolerance;
    unsigned maximum_steps;
    unsigned number_initialisations;
    int init_kind;
    const double* explicit_means;
};

struct mlpp_oracle_kmeans_result {
    double* centroids;
    unsigned* labels;
    double* step_seconds;   // entry 0 receives the whole-fit seconds
    double inertia;
    int converged;
    unsigned iterations;
};

static bool km_run(ml::Clustering::KMeans& km, const double* data, int64_t d, int64_t n, int64_t ld, unsigned k,
                   const mlpp_oracle_kmeans_options* opt, unsigned maximum_steps)
{
    if (opt->set_seed) km.set_seed(opt->seed);
    km.set_absolute_tolerance(opt->absolute_tolerance);
    km.set_maximum_steps(maximum_steps);
    km.set_number_initialisations(opt->number_initialisations);
    km.set_centroids_initialiser(make_initialiser(opt->init_kind, opt->explicit_means, d, k));
    return km.fit(wrap(data, d, n, ld));
}

// KMeans::fit; queries (d x n_queries) go to KMeans::assign_label: query_labels / query_sq receive the pairs.
int mlpp_ref_kmeans_fit(const double* data, int64_t d, int64_t n, int64_t ld, unsigned k, const mlpp_oracle_kmeans_options* opt,
                        mlpp_oracle_kmeans_result* out, int count_iterations, const double* queries, int64_t n_queries,
                        unsigned* query_labels, double* query_sq)
{
    try {
        ml::Clustering::KMeans km(k);
        const auto t0 = std::chrono::steady_clock::now();
        const bool converged = km_run(km, data, d, n, ld, k, opt, opt->maximum_steps);
        const double secs = seconds_since(t0);
        if (out->centroids) std::memcpy(out->centroids, km.centroids().data(), sizeof(double) * static_cast<size_t>(d) * k);
        if (out->labels) std::memcpy(out->labels, km.labels().data(), sizeof(unsigned) * static_cast<size_t>(n));
        if (out->step_seconds) out->step_seconds[0] = secs;
        out->inertia = km.inertia();
        out->converged = converged ? 1 : 0;
        out->iterations = 0;
        if (count_iterations) {
            unsigned iterations = opt->maximum_steps;
            if (converged && opt->number_initialisations == 1) {
                unsigned lo = 2, hi = opt->maximum_steps;   // smallest step limit that still converges
                while (lo < hi) {
                    const unsigned mid = lo + (hi - lo) / 2;
                    ml::Clustering::KMeans probe(k);
                    if (km_run(probe, data, d, n, ld, k, opt, mid)) hi = mid; else lo = mid + 1;
                }
                iterations = lo;
            }
            out->iterations = iterations;
        }
        for (int64_t q = 0; queries && q < n_queries; ++q) {
            Eigen::VectorXd x(d);
            for (int64_t i = 0; i < d; ++i) x[i] = queries[i + q * d];
            const auto r = km.assign_label(x);
            query_labels[q] = r.first;
            query_sq[q] = r.second;
        }
        return 0;
    } catch (const std::exception&) {
        return 1;
    }
}

// Forgy / RandomPartition / KPP on their own (Clustering.cpp:16-59).
void mlpp_ref_centroids_init(int kind, const double* data, int64_t d, int64_t n, int64_t ld, unsigned k, unsigned seed, int set_seed,
                             double* centroids)
{
    std::default_random_engine prng;
    if (set_seed) prng.seed(seed);
    Eigen::MatrixXd c(d, k);
    make_initialiser(kind, nullptr, d, k)->init(wrap(data, d, n, ld), prng, k, c);
    std::memcpy(centroids, c.data(), sizeof(double) * static_cast<size_t>(d) * k);
}

double mlpp_ref_xAx_symmetric(const double* A, int64_t dim, const double* x)
{
    Eigen::MatrixXd a(dim, dim);
    Eigen::VectorXd v(dim);
    std::memcpy(a.data(), A, sizeof(double) * static_cast<size_t>(dim * dim));
    std::memcpy(v.data(), x, sizeof(double) * static_cast<size_t>(dim));
    return ml::LinearAlgebra::xAx_symmetric(a, v);
}

void mlpp_ref_xxT(const double* x, int64_t dim, double* dest)
{
    Eigen::VectorXd v(dim);
    std::memcpy(v.data(), x, sizeof(double) * static_cast<size_t>(dim));
    Eigen::MatrixXd m;
    ml::LinearAlgebra::xxT(v, m);
    std::memcpy(dest, m.data(), sizeof(double) * static_cast<size_t>(dim * dim));
}

void mlpp_ref_add_a_xxT(const double* x, int64_t dim, double* dest, double a)
{
    Eigen::VectorXd v(dim);
    std::memcpy(v.data(), x, sizeof(double) * static_cast<size_t>(dim));
    Eigen::MatrixXd m(dim, dim);
    std::memcpy(m.data(), dest, sizeof(double) * static_cast<size_t>(dim * dim));
    ml::LinearAlgebra::add_a_xxT(v, m, a);
    std::memcpy(dest, m.data(), sizeof(double) * static_cast<size_t>(dim * dim));
}

}  // extern "C"
