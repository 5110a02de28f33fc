// Extension module `cppyml` (the reference's cppyml/cppyml.cpp:15-30).  Only the clustering
// submodule exists in this build: it is the hot path this repository accelerates; the reference's
// decision_trees / linear_regression / logistic_regression submodules are out of scope.
#include <pybind11/pybind11.h>

#include "ML/Version.hpp"

namespace py = pybind11;

void init_clustering(py::module_& m);

PYBIND11_MODULE(cppyml, m)
{
	m.doc() = "cppyml: Python bindings for the ML++ clustering algorithms on the B200 CUDA backend.";
	m.attr("__version__") = MLPP_VERSION;
	m.attr("backend") = MLPP_BACKEND;
	init_clustering(m);
}
