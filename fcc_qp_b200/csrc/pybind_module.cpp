// pybind11 module `fcc_qp_solver`: the reference's Python surface (src/main.cpp:19-56)
// over the B200 host class in include/fcc_qp.hpp.  Same class names, same attribute
// names (eps_bounds / eps_friction_cone renames included), same keyword arguments.
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <vector>

#include "fcc_qp.hpp"

namespace py = pybind11;

using fcc_qp::ConstMatrixView;
using fcc_qp::ConstVectorView;
using fcc_qp::FCCQP;
using fcc_qp::FCCQPDetails;
using fcc_qp::FCCQPOptions;
using fcc_qp::FCCQPSolution;

namespace {

// Any-layout float64 matrix (dtype conversions are copies, like pybind's Eigen caster;
// C- or F-ordered float64 arrays are consumed in place -- the reference's caster copies
// C-ordered ones, pybind11/eigen/matrix.h:516-533).
using Mat = py::array_t<double, py::array::forcecast>;
using Vec = py::array_t<double, py::array::c_style | py::array::forcecast>;

ConstMatrixView mat_view(const Mat& a, const char* name) {
  if (a.ndim() != 2) throw py::type_error(std::string(name) + " must be a 2-D array");
  return ConstMatrixView{a.data(), (int)a.shape(0), (int)a.shape(1),
                         (std::ptrdiff_t)(a.strides(0) / (py::ssize_t)sizeof(double)),
                         (std::ptrdiff_t)(a.strides(1) / (py::ssize_t)sizeof(double))};
}
ConstVectorView vec_view(const Vec& a, const char* name) {
  if (a.ndim() != 1) throw py::type_error(std::string(name) + " must be a 1-D array");
  return ConstVectorView{a.data(), (int)a.shape(0)};
}

}  // namespace

PYBIND11_MODULE(fcc_qp_solver, m) {
  m.doc() = "B200-native FCCQP solver with the Python surface of Brian-Acosta/fcc_qp";

  py::class_<FCCQPDetails>(m, "FCCQPDetails")
      .def_readwrite("n_iter", &FCCQPDetails::n_iter)
      .def_readwrite("eps_bounds", &FCCQPDetails::admm_residual_bounds)
      .def_readwrite("eps_friction_cone", &FCCQPDetails::admm_residual_friction_cone)
      .def_readwrite("bounds_viol", &FCCQPDetails::bounds_viol)
      .def_readwrite("friction_cone_viol", &FCCQPDetails::friction_cone_viol)
      .def_readwrite("solve_time", &FCCQPDetails::solve_time)
      .def_readwrite("factorization_time", &FCCQPDetails::factorization_time)
      // superset of the reference binding: status is computed there but not exposed
      .def_property_readonly("solve_status", [](const FCCQPDetails& d) { return (int)d.solve_status; });

  py::class_<FCCQPOptions>(m, "FCCQPOptions")
      .def(py::init<>())
      .def_readwrite("max_iter", &FCCQPOptions::max_iter)
      .def_readwrite("rho", &FCCQPOptions::rho)
      .def_readwrite("eps_fcone", &FCCQPOptions::eps_fcone)
      .def_readwrite("eps_bound", &FCCQPOptions::eps_bound)
      .def_readwrite("relaxation", &FCCQPOptions::relaxation);   // extension, default 1.0 = the reference

  py::class_<FCCQPSolution>(m, "FCCQPSolution")
      .def_readwrite("details", &FCCQPSolution::details)
      .def_property(
          "z",
          [](const FCCQPSolution& s) { return py::array_t<double>((py::ssize_t)s.z.size(), s.z.data()); },
          [](FCCQPSolution& s, const std::vector<double>& v) { s.z = v; });

  py::class_<FCCQP>(m, "FCCQP")
      .def(py::init<int, int, int, int, int>(), py::arg("num_vars"), py::arg("num_equality_constraints"),
           py::arg("nc"), py::arg("lambda_c_start"), py::arg("device") = 0)
      .def("set_rho", &FCCQP::set_rho)
      .def("set_max_iter", &FCCQP::set_max_iter)
      .def("set_warm_start", &FCCQP::set_warm_start)
      .def("set_options", &FCCQP::set_options)
      .def("contact_vars_start", &FCCQP::contact_vars_start)
      .def(
          "Solve",
          [](FCCQP& self, const Mat& Q, const Vec& b, const Mat& A_eq, const Vec& b_eq,
             const std::vector<double>& friction_coeffs, const Vec& lb, const Vec& ub) {
            const ConstMatrixView q = mat_view(Q, "Q");
            ConstMatrixView a = mat_view(A_eq, "A_eq");
            const ConstVectorView bv = vec_view(b, "b"), beq = vec_view(b_eq, "b_eq");
            const ConstVectorView l = vec_view(lb, "lb"), u = vec_view(ub, "ub");
            py::gil_scoped_release release;  // the reference holds the GIL for the whole Solve
            self.Solve(q, bv, a, beq, friction_coeffs, l, u);
          },
          py::arg("Q"), py::arg("b"), py::arg("A_eq"), py::arg("b_eq"), py::arg("friction_coeffs"),
          py::arg("lb"), py::arg("ub"))
      .def("GetSolution", &FCCQP::GetSolution)
      .def("GetWarmState",
           [](const FCCQP& self) {
             py::array_t<double> x(self.num_vars()), mx(self.num_vars()), mc(self.num_contact_vars());
             self.GetWarmState(x.mutable_data(), mx.mutable_data(), mc.mutable_data());
             return py::make_tuple(x, mx, mc);
           })
      .def("SetWarmState", [](FCCQP& self, const Vec& x, const Vec& mx, const Vec& mc) {
        if (x.size() != self.num_vars() || mx.size() != self.num_vars() || mc.size() != self.num_contact_vars())
          throw py::value_error("warm state arrays must be [num_vars], [num_vars], [nc]");
        self.SetWarmState(x.data(), mx.data(), mc.data());
      });
}
