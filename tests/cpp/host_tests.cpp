// CPU-side tests of the C++ host layer (no GPU needed): the initialisers' pseudo-random streams against
// the oracle, the LinearAlgebra identities of Tests/test_LinearAlgebra.cpp:8-75, the argument errors of
// the public API, and the N == K exact fits of Tests/test_EM.cpp:126-144 and Tests/test_KMeans.cpp:108-127.
// Links libML.so (product) and libmlpp_oracle.so (checker).  Prints "ok <n>" and exits 0 on success.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <limits>
#include <random>
#include <stdexcept>
#include <vector>

#include "ML/Clustering.hpp"
#include "ML/EM.hpp"
#include "ML/KMeans.hpp"
#include "ML/LinearAlgebra.hpp"
#include "src/Backend.hpp"

extern "C" {
void mlpp_oracle_centroids_init(int kind, const double* data, int64_t d, int64_t n, int64_t ld, unsigned k, unsigned seed, int set_seed, double* centroids);
double mlpp_oracle_xAx_symmetric(const double* A, int64_t dim, const double* x);
void mlpp_oracle_testdata_two_gaussians(double* data, unsigned* ground_truth);
}

static int checks = 0;
#define CHECK(cond) do { ++checks; if (!(cond)) { std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); std::exit(1); } } while (0)
#define CHECK_THROWS(expr, type) do { ++checks; bool caught = false; try { expr; } catch (const type&) { caught = true; } \
    if (!caught) { std::fprintf(stderr, "FAILED %s:%d: %s did not throw %s\n", __FILE__, __LINE__, #expr, #type); std::exit(1); } } while (0)

static void test_initialisers()
{
    Eigen::MatrixXd data(3, 400);
    mlpp_oracle_testdata_two_gaussians(data.data(), nullptr);
    const ml::Clustering::Forgy forgy;
    const ml::Clustering::RandomPartition random_partition;
    const ml::Clustering::KPP kpp;
    const ml::Clustering::CentroidsInitialiser* initialisers[] = {&forgy, &random_partition, &kpp};
    for (int kind = 0; kind < 3; ++kind) {
        for (unsigned k : {1u, 2u, 5u}) {
            std::default_random_engine prng;
            prng.seed(63413131);
            Eigen::MatrixXd got(3, k), want(3, k);
            initialisers[kind]->init(data, prng, k, got);
            mlpp_oracle_centroids_init(kind, data.data(), 3, 400, 3, k, 63413131, 1, want.data());
            for (Eigen::Index e = 0; e < got.size(); ++e) CHECK(got.data()[e] == want.data()[e]);   // bit for bit
        }
    }
    // ClosestCentroid: exactly one unit responsibility per row, on the nearest centroid
    std::default_random_engine prng, twin;
    Eigen::MatrixXd resp(400, 3), centroids(3, 3);
    const ml::Clustering::ClosestCentroid closest(std::make_shared<ml::Clustering::Forgy>());
    closest.init(data, prng, 3, resp);
    forgy.init(data, twin, 3, centroids);
    for (Eigen::Index i = 0; i < 400; ++i) {
        double row = 0, best = std::numeric_limits<double>::infinity();
        int arg = -1;
        for (int k = 0; k < 3; ++k) {
            row += resp(i, k);
            double d2 = 0;
            for (int l = 0; l < 3; ++l) d2 += (data(l, i) - centroids(l, k)) * (data(l, i) - centroids(l, k));
            if (d2 < best) { best = d2; arg = k; }
        }
        CHECK(row == 1.0);
        CHECK(resp(i, arg) == 1.0);
    }
    CHECK_THROWS(ml::Clustering::ClosestCentroid(nullptr), std::invalid_argument);
}

// detail::draw_discrete against the library call it stands for (Clustering.cpp:55-56): the same index and the same
// generator state afterwards, for weight vectors shaped like KPP's (zeros at the chosen points, wide dynamic range).
static void test_draw_discrete()
{
    std::mt19937_64 shapes(12345);
    for (int trial = 0; trial < 400; ++trial) {
        const size_t n = trial < 4 ? static_cast<size_t>(trial + 1) : 2 + shapes() % 5000;
        std::vector<double> weights(n);
        for (double& w : weights) {
            const double u = static_cast<double>(shapes() >> 11) / 9007199254740992.0;
            w = (shapes() % 7 == 0) ? 0.0 : std::exp(20.0 * u - 10.0);
        }
        if (trial % 5 == 0) std::fill(weights.begin(), weights.end(), 1.0);   // the first KPP draw
        weights[shapes() % n] += 1e-3;   // never all zero
        std::default_random_engine a, b;
        a.seed(1000 + trial);
        b.seed(1000 + trial);
        for (int draw = 0; draw < 3; ++draw) {
            std::discrete_distribution<Eigen::Index> library(weights.begin(), weights.end());
            const Eigen::Index want = library(a);
            const Eigen::Index got = ml::detail::draw_discrete(weights, b);
            CHECK(got == want);
            CHECK(a == b);
        }
    }
}

static void test_linear_algebra()
{
    std::default_random_engine rng(7);
    std::normal_distribution<double> normal;
    for (Eigen::Index n : {4, 20}) {
        Eigen::MatrixXd A(n, n);
        Eigen::VectorXd x(n);
        for (Eigen::Index i = 0; i < n; ++i) {
            x[i] = normal(rng);
            for (Eigen::Index j = 0; j <= i; ++j) A(i, j) = A(j, i) = normal(rng);
        }
        double direct = 0;
        for (Eigen::Index i = 0; i < n; ++i)
            for (Eigen::Index j = 0; j < n; ++j) direct += x[i] * A(i, j) * x[j];
        const double got = ml::LinearAlgebra::xAx_symmetric(A, x);
        CHECK(std::abs(got - direct) <= 1e-13 * std::max(1.0, std::abs(direct)));
        CHECK(std::abs(got - mlpp_oracle_xAx_symmetric(A.data(), n, x.data())) <= 1e-13 * std::max(1.0, std::abs(direct)));
        Eigen::MatrixXd outer;
        ml::LinearAlgebra::xxT(x, outer);
        Eigen::MatrixXd acc = Eigen::MatrixXd::Zero(n, n);
        ml::LinearAlgebra::add_a_xxT(x, acc, -0.3);
        for (Eigen::Index i = 0; i < n; ++i)
            for (Eigen::Index j = 0; j < n; ++j) {
                CHECK(outer(i, j) == x[i] * x[j]);
                CHECK(std::abs(acc(i, j) + 0.3 * x[i] * x[j]) <= 1e-15 * std::abs(x[i] * x[j]) + 1e-300);
            }
    }
    Eigen::MatrixXd rect(3, 2), square(3, 3);
    Eigen::VectorXd x3(3), x2(2);
    CHECK_THROWS(ml::LinearAlgebra::xAx_symmetric(rect, x3), std::invalid_argument);
    CHECK_THROWS(ml::LinearAlgebra::xAx_symmetric(square, x2), std::invalid_argument);
    CHECK_THROWS(ml::LinearAlgebra::add_a_xxT(x3, rect, 1.0), std::invalid_argument);
    CHECK_THROWS(ml::LinearAlgebra::add_a_xxT(x2, square, 1.0), std::invalid_argument);
}

static void test_argument_errors()
{
    CHECK_THROWS(ml::EM(0), std::invalid_argument);
    CHECK_THROWS(ml::Clustering::KMeans(0), std::invalid_argument);
    ml::EM em(2);
    CHECK_THROWS(em.set_absolute_tolerance(-1e-3), std::domain_error);
    CHECK_THROWS(em.set_relative_tolerance(-1e-3), std::domain_error);
    CHECK_THROWS(em.set_maximum_steps(1), std::invalid_argument);
    CHECK_THROWS(em.set_means_initialiser(nullptr), std::invalid_argument);
    CHECK_THROWS(em.set_responsibilities_initialiser(nullptr), std::invalid_argument);
    CHECK_THROWS(em.covariance(2), std::invalid_argument);
    Eigen::MatrixXd one_point(3, 1), no_rows(0, 5);
    CHECK_THROWS(em.fit(one_point), std::invalid_argument);
    CHECK_THROWS(em.fit(no_rows), std::invalid_argument);
    ml::Clustering::KMeans km(2);
    CHECK_THROWS(km.set_absolute_tolerance(-1.0), std::domain_error);
    CHECK_THROWS(km.set_maximum_steps(1), std::invalid_argument);
    CHECK_THROWS(km.set_number_initialisations(0), std::invalid_argument);
    CHECK_THROWS(km.set_centroids_initialiser(nullptr), std::invalid_argument);
    CHECK_THROWS(km.fit(one_point), std::invalid_argument);
    CHECK(em.number_components() == 2 && em.number_clusters() == 2 && km.number_clusters() == 2);
    CHECK(!em.converged() && !km.converged());
}

static void test_deterministic_fits()
{
    // as many components as points: exact fit without touching the device
    Eigen::MatrixXd data(2, 3);
    data << 0.5, 0.3, 0.1,
            0.1, 0.2, -0.4;
    ml::EM em(3);
    CHECK(em.fit(data));
    CHECK(em.converged());
    CHECK(em.labels().size() == 3);
    for (unsigned i = 0; i < 3; ++i) {
        CHECK(em.labels()[i] == i);
        for (Eigen::Index l = 0; l < 2; ++l) CHECK(em.means()(l, i) == data(l, i));
        CHECK(em.covariance(i).rows() == 2 && em.covariance(i)(0, 0) == 0.0 && em.covariance(i)(1, 0) == 0.0);
        for (unsigned j = 0; j < 3; ++j) CHECK(em.responsibilities()(i, j) == (i == j ? 1.0 : 0.0));
    }
    CHECK(std::isinf(em.log_likelihood()) && em.log_likelihood() > 0);
    ml::Clustering::KMeans km(3);
    CHECK(km.fit(data));
    CHECK(km.converged() && km.inertia() == 0.0);
    for (unsigned i = 0; i < 3; ++i) {
        CHECK(km.labels()[i] == i);
        Eigen::VectorXd x(2);
        x[0] = data(0, i); x[1] = data(1, i);
        const auto label = km.assign_label(x);
        CHECK(label.first == i && label.second == 0.0);
    }
    const ml::Clustering::Model& as_model = km;
    CHECK(as_model.centroids().cols() == 3 && as_model.number_clusters() == 3);
}

int main()
{
    test_initialisers();
    test_draw_discrete();
    test_linear_algebra();
    test_argument_errors();
    test_deterministic_fits();
    std::printf("ok %d\n", checks);
    return 0;
}
