// ORACLE — TEST INFRASTRUCTURE ONLY (see gtn_cpu.h).  pybind11 bindings that
// expose the CPU restatement of the GTN ops under the Python names the
// reference uses (SURVEY.md §8(b) Level 3), once for float ("f32", GTN's own
// scalar type) and once for double ("f64", the truth build).
#include <pybind11/functional.h>
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <atomic>
#include <cstdlib>
#include <thread>

#include "gtn_cpu.h"

namespace py = pybind11;
using namespace og;

namespace {

unsigned poolSize() {
  if (const char* e = std::getenv("GTN_ORACLE_THREADS")) {
    int v = std::atoi(e);
    if (v > 0) return (unsigned)v;
  }
  unsigned hw = std::thread::hardware_concurrency();
  return hw ? hw : 1;
}

// gtn.parallel_for(fn, ints): runs fn(i) on a pool of hardware-concurrency
// threads; each task takes the GIL for the Python callback while the bound
// graph ops release it (ctc.py:65,83; asg.py:129,170,236; stc.py:100,118;
// transducer.py:232,296,327,506,541).  The first exception is re-raised.
void parallelFor(py::function fn, py::iterable items) {
  std::vector<py::object> args;
  for (auto h : items) args.push_back(py::reinterpret_borrow<py::object>(h));
  const size_t n = args.size();
  if (n == 0) return;
  unsigned nt = (unsigned)std::min<size_t>(poolSize(), n);
  std::atomic<size_t> next{0};
  std::exception_ptr err;
  std::mutex errMu;
  {
    py::gil_scoped_release rel;
    auto work = [&]() {
      while (true) {
        size_t i = next.fetch_add(1);
        if (i >= n) break;
        py::gil_scoped_acquire acq;
        try {
          fn(args[i]);
        } catch (...) {
          std::lock_guard<std::mutex> lk(errMu);
          if (!err) err = std::current_exception();
        }
      }
    };
    if (nt == 1) {
      work();
    } else {
      std::vector<std::thread> th;
      for (unsigned t = 0; t < nt; ++t) th.emplace_back(work);
      for (auto& t : th) t.join();
    }
  }
  if (err) std::rethrow_exception(err);
}

template <typename R>
void bindAll(py::module_& m) {
  using G = GraphT<R>;
  using NoGil = py::call_guard<py::gil_scoped_release>;

  py::class_<G>(m, "Graph")
      .def(py::init([](py::object cg) { return G(PyObject_IsTrue(cg.ptr()) == 1); }),
           py::arg("calc_grad") = true)
      .def("add_node", &G::addNode, py::arg("start") = false, py::arg("accept") = false)
      .def("add_arc", [](G& g, int s, int d, int l) { return g.addArc(s, d, l); },
           py::arg("src_node"), py::arg("dst_node"), py::arg("label"))
      .def("add_arc",
           [](G& g, int s, int d, int il, int ol, double w) { return g.addArc(s, d, il, ol, (R)w); },
           py::arg("src_node"), py::arg("dst_node"), py::arg("ilabel"), py::arg("olabel"),
           py::arg("weight") = 0.0)
      .def("make_accept", &G::makeAccept)
      .def("arc_sort", &G::arcSort, py::arg("olabel") = false, NoGil())
      .def("mark_arc_sorted", &G::markArcSorted, py::arg("olabel") = false)
      .def("ilabel_sorted", &G::ilabelSorted)
      .def("olabel_sorted", &G::olabelSorted)
      .def("num_nodes", &G::numNodes)
      .def("num_arcs", &G::numArcs)
      .def("num_start", [](const G& g) { return (int)g.start().size(); })
      .def("num_accept", [](const G& g) { return (int)g.accept().size(); })
      .def("start", [](const G& g) { return g.start(); })
      .def("accept", [](const G& g) { return g.accept(); })
      .def("is_start", &G::isStart)
      .def("is_accept", &G::isAccept)
      .def("in_arcs", [](const G& g, int n) { return g.in(n); })
      .def("out_arcs", [](const G& g, int n) { return g.out(n); })
      .def("src_node", &G::srcNode)
      .def("dst_node", &G::dstNode)
      .def("ilabel", &G::ilabel)
      .def("olabel", &G::olabel)
      .def("weight", &G::weight)
      .def("set_weight", &G::setWeight)
      .def("item", &G::item)
      .def_property("calc_grad", &G::calcGrad, &G::setCalcGrad)
      .def("grad", [](G& g) { return g.grad(); })
      .def("has_grad", &G::hasGrad)
      .def("zero_grad", &G::zeroGrad)
      // set_weights(int) reads num_arcs float32 values from a raw host
      // pointer (what the reference passes: tensor.data_ptr(), ctc.py:44);
      // sequences / arrays are copied element-wise.
      .def("set_weights",
           [](G& g, py::object w) {
             if (py::isinstance<py::int_>(w)) {
               auto p = reinterpret_cast<const float*>(w.cast<std::uintptr_t>());
               py::gil_scoped_release rel;
               g.setWeightsF32(p);
               return;
             }
             auto arr = py::array_t<double, py::array::c_style | py::array::forcecast>::ensure(w);
             if (!arr || (int)arr.size() != g.numArcs())
               throw std::invalid_argument("[set_weights] need num_arcs weights");
             g.setWeightsF64(arr.data());
           })
      .def("weights_to_numpy",
           [](const G& g) {
             py::array_t<R> out(g.numArcs());
             std::memcpy(out.mutable_data(), g.weights().data(), sizeof(R) * g.numArcs());
             return out;
           })
      .def("weights_to_list", [](const G& g) { return g.weights(); })
      .def("labels_to_list", &G::labels, py::arg("ilabel") = true)
      // bulk structure dump for index-exact comparisons in tests
      .def("arcs_to_numpy",
           [](const G& g) {
             const auto& t = g.topo();
             auto mk = [](const std::vector<int>& v) {
               py::array_t<int> a(v.size());
               if (!v.empty()) std::memcpy(a.mutable_data(), v.data(), sizeof(int) * v.size());
               return a;
             };
             return py::make_tuple(mk(t.src), mk(t.dst), mk(t.il), mk(t.ol));
           })
      .def("in_order",
           [](const G& g) {
             std::vector<int> v;
             for (int n = 0; n < g.numNodes(); ++n)
               for (int a : g.in(n)) v.push_back(a);
             return v;
           })
      .def("out_order", [](const G& g) {
        std::vector<int> v;
        for (int n = 0; n < g.numNodes(); ++n)
          for (int a : g.out(n)) v.push_back(a);
        return v;
      });

  m.def("linear_graph_", &linearGraph<R>, py::arg("M"), py::arg("N"),
        py::arg("calc_grad") = true, NoGil());
  m.def("scalar_graph", &scalarGraph<R>, py::arg("weight"), py::arg("calc_grad") = true);
  m.def("negate", &negate<R>, NoGil());
  m.def("add", [](const G& a, const G& b) { return addOrSub(a, b, false); }, NoGil());
  m.def("subtract", [](const G& a, const G& b) { return addOrSub(a, b, true); }, NoGil());
  m.def("project_input", [](const G& g) { return project(g, true); }, NoGil());
  m.def("project_output", [](const G& g) { return project(g, false); }, NoGil());
  m.def("remove", [](const G& g, int label) { return removeLabel(g, label, label); },
        py::arg("g"), py::arg("label") = kEps, NoGil());
  m.def("remove", [](const G& g, int il, int ol) { return removeLabel(g, il, ol); },
        py::arg("g"), py::arg("ilabel"), py::arg("olabel"), NoGil());
  m.def("compose", &compose<R>, NoGil());
  m.def("intersect", &compose<R>, NoGil());
  m.def("forward_score", [](const G& g) { return shortestDistance(g, false); }, NoGil());
  m.def("viterbi_score", [](const G& g) { return shortestDistance(g, true); }, NoGil());
  m.def("viterbi_path", &viterbiPath<R>, NoGil());
  m.def("backward", [](G g, bool retain) { backward(g, retain); }, py::arg("g"),
        py::arg("retain_graph") = false, NoGil());
  m.def("backward", [](G g, const G& seed, bool retain) { backward(g, seed, retain); },
        py::arg("g"), py::arg("grad"), py::arg("retain_graph") = false, NoGil());
  m.def("equal", &equal<R>);
  m.def("isomorphic", &isomorphic<R>);
  m.def("clone", [](const G& g) { return G::deepCopy(g); });
  m.def("loadtxt", [](const std::string& path) {
    std::ifstream in(path);
    if (!in) throw std::invalid_argument("[loadtxt] cannot open " + path);
    return loadTxt<R>(in);
  });
  m.def("savetxt", [](const std::string& path, const G& g) {
    std::ofstream out(path);
    if (!out) throw std::invalid_argument("[savetxt] cannot open " + path);
    saveTxt(out, g);
  });
  m.def("load", [](const std::string& path) {
    std::ifstream in(path, std::ios::binary);
    if (!in) throw std::invalid_argument("[load] cannot open " + path);
    return loadBin<R>(in);
  });
  m.def("save", [](const std::string& path, const G& g) {
    std::ofstream out(path, std::ios::binary);
    if (!out) throw std::invalid_argument("[save] cannot open " + path);
    saveBin(out, g);
  });
  m.def("parallel_for", &parallelFor);
  m.attr("epsilon") = kEps;
}

}  // namespace

PYBIND11_MODULE(_gtn_oracle, m) {
  m.doc() = "CPU restatement of the GTN ops used by gtn_applications (oracle; tests only)";
  auto f32 = m.def_submodule("f32");
  auto f64 = m.def_submodule("f64");
  bindAll<float>(f32);
  bindAll<double>(f64);
  m.def("pool_size", &poolSize);
}
