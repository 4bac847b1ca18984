// bind_core.cpp -- `rgpot_b200._core`: the compiled Python surface of the reference's nanobind module
// `rgpot._core` (python/bind/bind_module.cpp:84-96 evaluate_lj, :160-173 LJPot.__call__, :142-147
// __version__ / has_metatomic_dlopen), on the CUDA plugin classes of cuda_pots.hpp instead of the CPU
// potentials.  pybind11 is what this image offers (nanobind is absent); names, argument order, array
// types (C-contiguous float64 / int32 numpy arrays on the host) and the error behaviour are the
// reference's: shape errors are std::invalid_argument, i.e. Python ValueError; outputs are fresh numpy
// arrays that own their memory.  Beyond the reference's surface (which exposes LJ only):
// CuH2Pot, evaluate_cuh2 and evaluate_batch (ForceBatch semantics, one engine call).
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <array>
#include <cstring>
#include <stdexcept>
#include <vector>

#include "cuda_pots.hpp"

namespace py = pybind11;
using NpF64 = py::array_t<double, py::array::c_style | py::array::forcecast>;
using NpI32 = py::array_t<int32_t, py::array::c_style | py::array::forcecast>;

namespace {

struct System {
  std::vector<double> pos;
  std::vector<int> types;
  double box[9];
  size_t n = 0;
};

System marshal(const NpF64 &positions, const NpI32 &atom_types, const NpF64 &box) {
  if (positions.ndim() != 2 || positions.shape(1) != 3)
    throw std::invalid_argument("positions must have shape (n_atoms, 3)");
  System s;
  s.n = static_cast<size_t>(positions.shape(0));
  if (atom_types.ndim() != 1 || static_cast<size_t>(atom_types.shape(0)) != s.n)
    throw std::invalid_argument("atom_types length must match n_atoms");
  if (box.ndim() != 2 || box.shape(0) != 3 || box.shape(1) != 3)
    throw std::invalid_argument("box must have shape (3, 3)");
  s.pos.assign(positions.data(), positions.data() + 3 * s.n);
  s.types.assign(atom_types.data(), atom_types.data() + s.n);
  std::memcpy(s.box, box.data(), sizeof s.box);
  return s;
}

template <class Pot>
py::tuple evaluate(const Pot &pot, const NpF64 &positions, const NpI32 &atom_types, const NpF64 &box) {
  System s = marshal(positions, atom_types, box);
  py::array_t<double> forces({static_cast<py::ssize_t>(s.n), static_cast<py::ssize_t>(3)});
  rgpot::ForceInput in{s.n, s.pos.data(), s.types.data(), s.box};
  rgpot::ForceOut out{forces.mutable_data(), 0.0, 0.0};
  pot.forceImpl(in, &out);
  return py::make_tuple(out.energy, forces, out.variance);
}

py::tuple evaluate_lj(const NpF64 &positions, const NpI32 &atom_types, const NpF64 &box) {
  if (positions.ndim() != 2 || atom_types.ndim() != 1 || box.ndim() != 2)
    throw std::invalid_argument("evaluate_lj expects positions (n,3), atom_types (n,), box (3,3)");
  static rgpot::LJCudaPot pot;  // default LJConfig, like `rgpot::LJPot pot;` upstream
  return evaluate(pot, positions, atom_types, box);
}

py::tuple evaluate_cuh2(const NpF64 &positions, const NpI32 &atom_types, const NpF64 &box) {
  static rgpot::CuH2CudaPot pot;
  return evaluate(pot, positions, atom_types, box);
}

// [(positions, atom_types, box), ...] -> [(energy, forces, variance), ...] in ONE forceBatch call
template <class Pot>
py::list evaluate_batch(const Pot &pot, const py::list &systems) {
  const size_t n = systems.size();
  std::vector<System> sys;
  sys.reserve(n);
  for (const auto &item : systems) {
    const py::tuple t = item.cast<py::tuple>();
    if (t.size() != 3) throw std::invalid_argument("each system is (positions, atom_types, box)");
    sys.push_back(marshal(t[0].cast<NpF64>(), t[1].cast<NpI32>(), t[2].cast<NpF64>()));
  }
  std::vector<py::array_t<double>> forces;
  std::vector<rgpot::ForceInput> in;  // (ForceInput has const members: built in place, never assigned)
  std::vector<rgpot::ForceOut> out;
  in.reserve(n);
  out.reserve(n);
  forces.reserve(n);
  for (size_t i = 0; i < n; ++i) {
    forces.emplace_back(std::vector<py::ssize_t>{static_cast<py::ssize_t>(sys[i].n), 3});
    in.push_back(rgpot::ForceInput{sys[i].n, sys[i].pos.data(), sys[i].types.data(), sys[i].box});
    out.push_back(rgpot::ForceOut{forces[i].mutable_data(), 0.0, 0.0});
  }
  rgpot::ForceBatch batch{n, in.data(), out.data()};
  pot.forceBatchImpl(batch);
  py::list res;
  for (size_t i = 0; i < n; ++i) res.append(py::make_tuple(out[i].energy, forces[i], out[i].variance));
  return res;
}

} // namespace

PYBIND11_MODULE(_core, m) {
  m.doc() = "rgpot_b200 compiled Python surface (pybind11): the reference's rgpot._core names on the B200 engine.";
  m.attr("__version__") = "2.5.2+b200";
  m.attr("has_metatomic_dlopen") = false;

  m.def("evaluate_lj", &evaluate_lj, py::arg("positions"), py::arg("atom_types"), py::arg("box"));
  m.def("evaluate_cuh2", &evaluate_cuh2, py::arg("positions"), py::arg("atom_types"), py::arg("box"));

  py::class_<rgpot::LJCudaPot>(m, "LJPot")
      .def(py::init<>())
      .def(py::init([](double u0, double cutoff, double psi, int device) {
             rgpot::LJCudaConfig c;
             c.u0 = u0;
             c.cutoff = cutoff;
             c.psi = psi;
             c.device = device;
             return new rgpot::LJCudaPot(c);
           }),
           py::arg("u0"), py::arg("cutoff"), py::arg("psi"), py::arg("device") = -1)
      .def("__call__",
           [](rgpot::LJCudaPot &self, const NpF64 &p, const NpI32 &t, const NpF64 &b) { return evaluate(self, p, t, b); },
           py::arg("positions"), py::arg("atom_types"), py::arg("box"))
      .def("forceBatch", [](rgpot::LJCudaPot &self, const py::list &systems) { return evaluate_batch(self, systems); });

  py::class_<rgpot::CuH2CudaPot>(m, "CuH2Pot")
      .def(py::init<>())
      .def("__call__",
           [](rgpot::CuH2CudaPot &self, const NpF64 &p, const NpI32 &t, const NpF64 &b) { return evaluate(self, p, t, b); },
           py::arg("positions"), py::arg("atom_types"), py::arg("box"))
      .def("forceBatch", [](rgpot::CuH2CudaPot &self, const py::list &systems) { return evaluate_batch(self, systems); });
}
