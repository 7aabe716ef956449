// ORACLE — TEST INFRASTRUCTURE ONLY.  Nothing under gtn_applications_b200/ may
// include, link or import this file.  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs use it (as the checker / the
// timed CPU baseline, never as the product path).
//
// What this is: a CPU restatement, written from scratch, of the subset of the
// external GTN C++ library (facebookresearch/gtn, NOT vendored in
// /root/reference; version unpinned by the reference: requirements.txt:1 pins
// only editdistance, README.md:11 says "build python bindings") that the
// reference's hot path calls:
//   criterions/ctc.py:40-51,78-80   linear_graph / set_weights / intersect /
//                                   forward_score / negate / backward / grad
//   criterions/asg.py:53-115,158    + subtract, mark_arc_sorted
//   criterions/stc.py:22-86,113     + compose, weighted add_arc
//   criterions/transducer.py:15-123,199-348,461-556
//                                   + remove, project_input/output,
//                                     viterbi_score/path, seeded backward
//   utils.py:261 (gtn.load), tests/transducer_test.py:535 (gtn.loadtxt)
// Semantics follow GTN's published algorithm:
// explicit composition (reverse co-reachability BFS, then forward BFS build
// in matcher order with (arc1, arc2) provenance), Kahn-order shortest
// distance in the log / tropical semiring, and a reverse-mode tape.
//
// PARITY PIN: numerical results are pinned by the reference's own golden
// tests (tests/gtn_ctc_test.py:24-80, tests/gtn_asg_test.py:25-124,
// tests/gtn_stc_test.py:25-51, tests/transducer_test.py:100-566), which
// tests/test_oracle_reference.py runs UNCHANGED against this shim in the
// build container.  Composed-lattice node/arc numbering is pinned by no
// reference test — for those indices parity is "unpinned".
//
// The scalar type R is float (GTN's type; module `gtn`) or double (the
// "truth" build used to judge both the fp32 oracle and the CUDA path).
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <functional>
#include <limits>
#include <memory>
#include <mutex>
#include <queue>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <unordered_set>
#include <utility>
#include <vector>

namespace og {

constexpr int kEps = -1;

template <typename R>
class GraphT {
 public:
  using GradFunc = std::function<void(std::vector<GraphT>&, GraphT&)>;

  struct Node {
    bool start{false};
    bool accept{false};
    std::vector<int> in;
    std::vector<int> out;
  };
  struct Topo {
    std::vector<Node> nodes;
    std::vector<int> src, dst, il, ol;
    std::vector<int> start, accept;
    bool ilabelSorted{false};
    bool olabelSorted{false};
  };
  struct Auto {
    GradFunc func;
    std::vector<GraphT> inputs;
    std::unique_ptr<GraphT> grad;
    bool calcGrad{true};
    std::mutex mu;
  };

  explicit GraphT(bool calcGrad = true)
      : t_(std::make_shared<Topo>()),
        w_(std::make_shared<std::vector<R>>()),
        a_(std::make_shared<Auto>()) {
    a_->calcGrad = calcGrad;
  }

  GraphT(GradFunc f, std::vector<GraphT> inputs) : GraphT(false) {
    bool cg = false;
    for (auto& g : inputs) cg = cg || g.calcGrad();
    a_->calcGrad = cg;
    if (cg) {
      a_->func = std::move(f);
      a_->inputs = std::move(inputs);
    }
  }

  // ---- construction -------------------------------------------------------
  int addNode(bool start = false, bool accept = false) {
    int idx = numNodes();
    t_->nodes.emplace_back();
    t_->nodes.back().start = start;
    t_->nodes.back().accept = accept;
    if (start) t_->start.push_back(idx);
    if (accept) t_->accept.push_back(idx);
    return idx;
  }
  int addArc(int s, int d, int label) { return addArc(s, d, label, label, R(0)); }
  int addArc(int s, int d, int il, int ol, R w = R(0)) {
    if (s < 0 || s >= numNodes() || d < 0 || d >= numNodes())
      throw std::out_of_range("[Graph::addArc] invalid node index");
    int idx = numArcs();
    t_->src.push_back(s);
    t_->dst.push_back(d);
    t_->il.push_back(il);
    t_->ol.push_back(ol);
    w_->push_back(w);
    t_->nodes[s].out.push_back(idx);
    t_->nodes[d].in.push_back(idx);
    t_->ilabelSorted = false;
    t_->olabelSorted = false;
    return idx;
  }
  void makeAccept(int n) {
    if (!t_->nodes[n].accept) {
      t_->nodes[n].accept = true;
      t_->accept.push_back(n);
    }
  }

  // ---- accessors ----------------------------------------------------------
  int numNodes() const { return (int)t_->nodes.size(); }
  int numArcs() const { return (int)t_->src.size(); }
  const std::vector<int>& start() const { return t_->start; }
  const std::vector<int>& accept() const { return t_->accept; }
  bool isStart(int n) const { return t_->nodes[n].start; }
  bool isAccept(int n) const { return t_->nodes[n].accept; }
  const std::vector<int>& in(int n) const { return t_->nodes[n].in; }
  const std::vector<int>& out(int n) const { return t_->nodes[n].out; }
  int srcNode(int a) const { return t_->src[a]; }
  int dstNode(int a) const { return t_->dst[a]; }
  int ilabel(int a) const { return t_->il[a]; }
  int olabel(int a) const { return t_->ol[a]; }
  R weight(int a) const { return (*w_)[a]; }
  void setWeight(int a, R v) { (*w_)[a] = v; }
  std::vector<R>& weights() { return *w_; }
  const std::vector<R>& weights() const { return *w_; }
  bool ilabelSorted() const { return t_->ilabelSorted; }
  bool olabelSorted() const { return t_->olabelSorted; }
  const Topo& topo() const { return *t_; }
  std::uintptr_t id() const { return reinterpret_cast<std::uintptr_t>(a_.get()); }

  R item() const {
    if (numArcs() != 1)
      throw std::invalid_argument("[Graph::item] graph must have exactly one arc");
    return (*w_)[0];
  }

  // arc_sort(olabel): sorts every node's in/out arc lists by that label; arc
  // numbers do not change.  (GTN uses std::sort; the order among equal labels
  // is therefore unspecified there — stable order is used here.)
  void arcSort(bool olabel = false) {
    if ((olabel && t_->olabelSorted) || (!olabel && t_->ilabelSorted)) return;
    const auto& lab = olabel ? t_->ol : t_->il;
    auto cmp = [&lab](int a, int b) { return lab[a] < lab[b]; };
    for (auto& n : t_->nodes) {
      std::stable_sort(n.in.begin(), n.in.end(), cmp);
      std::stable_sort(n.out.begin(), n.out.end(), cmp);
    }
    t_->ilabelSorted = !olabel;
    t_->olabelSorted = olabel;
  }
  void markArcSorted(bool olabel = false) {
    if (olabel) t_->olabelSorted = true;
    else t_->ilabelSorted = true;
  }

  void setWeightsF32(const float* p) {
    for (int i = 0; i < numArcs(); ++i) (*w_)[i] = (R)p[i];
  }
  void setWeightsF64(const double* p) {
    for (int i = 0; i < numArcs(); ++i) (*w_)[i] = (R)p[i];
  }
  std::vector<int> labels(bool ilabel = true) const { return ilabel ? t_->il : t_->ol; }

  // ---- autograd -----------------------------------------------------------
  bool calcGrad() const { return a_->calcGrad; }
  void setCalcGrad(bool v) {
    a_->calcGrad = v;
    if (!v) {
      a_->func = nullptr;
      a_->inputs.clear();
      a_->grad.reset();
    }
  }
  bool hasGrad() const { return a_->grad != nullptr; }
  GraphT& grad() {
    if (!a_->grad) throw std::logic_error("[Graph::grad] gradient not calculated yet");
    return *a_->grad;
  }
  void zeroGrad() { a_->grad.reset(); }
  GradFunc& gradFunc() { return a_->func; }
  std::vector<GraphT>& inputs() { return a_->inputs; }

  // Gradients accumulate when a graph feeds several ops (asg.py:111-114 uses
  // g_emissions and g_transitions twice; transducer.py shares one transitions
  // graph across all utterances of the batch, hence the mutex).
  void addGrad(std::vector<R>&& g) {
    if (!calcGrad()) return;
    if ((int)g.size() != numArcs())
      throw std::logic_error("[Graph::addGrad] size mismatch");
    std::lock_guard<std::mutex> lk(a_->mu);
    if (!a_->grad) {
      a_->grad.reset(new GraphT(false));
      a_->grad->t_ = t_;
      *a_->grad->w_ = std::move(g);
    } else {
      auto& acc = *a_->grad->w_;
      for (size_t i = 0; i < g.size(); ++i) acc[i] += g[i];
    }
  }
  void addGrad(const GraphT& other) { addGrad(std::vector<R>(other.weights())); }

  // deep copy of topology and weights, detached from the tape
  static GraphT deepCopy(const GraphT& g) {
    GraphT out(g.calcGrad());
    *out.t_ = *g.t_;
    *out.w_ = *g.w_;
    return out;
  }

 private:
  std::shared_ptr<Topo> t_;
  std::shared_ptr<std::vector<R>> w_;
  std::shared_ptr<Auto> a_;
};

// ---------------------------------------------------------------------------
// creations
// ---------------------------------------------------------------------------
// linear_graph(M, N): nodes 0..M (0 start, M accept); arc m*N+n goes m->m+1
// with label n; marked sorted on both labels.  (ctc.py:40, asg.py:96,
// stc.py:74, transducer.py:262,489)
template <typename R>
GraphT<R> linearGraph(int M, int N, bool calcGrad = true) {
  GraphT<R> g(calcGrad);
  g.addNode(true, M == 0);
  for (int m = 1; m <= M; ++m) {
    g.addNode(false, m == M);
    for (int n = 0; n < N; ++n) g.addArc(m - 1, m, n);
  }
  g.markArcSorted(false);
  g.markArcSorted(true);
  return g;
}

template <typename R>
GraphT<R> scalarGraph(R w, bool calcGrad = true) {
  GraphT<R> g(calcGrad);
  g.addNode(true);
  g.addNode(false, true);
  g.addArc(0, 1, 0, 0, w);
  return g;
}

// ---------------------------------------------------------------------------
// scalar-graph arithmetic on the tape (ctc.py:49, asg.py:115, transducer.py:288-290)
// ---------------------------------------------------------------------------
template <typename R>
GraphT<R> negate(const GraphT<R>& g) {
  if (g.numArcs() != 1) throw std::logic_error("[negate] input must have only one arc");
  auto gf = [](std::vector<GraphT<R>>& in, GraphT<R>& d) {
    in[0].addGrad(std::vector<R>{-d.item()});
  };
  GraphT<R> out(gf, {g});
  out.addNode(true);
  out.addNode(false, true);
  out.addArc(0, 1, 0, 0, -g.item());
  return out;
}

template <typename R>
GraphT<R> addOrSub(const GraphT<R>& a, const GraphT<R>& b, bool sub) {
  if (a.numArcs() != 1 || b.numArcs() != 1)
    throw std::logic_error("[add/subtract] inputs must have only one arc");
  auto gf = [sub](std::vector<GraphT<R>>& in, GraphT<R>& d) {
    in[0].addGrad(std::vector<R>{d.item()});
    in[1].addGrad(std::vector<R>{sub ? -d.item() : d.item()});
  };
  GraphT<R> out(gf, {a, b});
  out.addNode(true);
  out.addNode(false, true);
  out.addArc(0, 1, 0, 0, sub ? a.item() - b.item() : a.item() + b.item());
  return out;
}

// project_input / project_output: copy with olabel:=ilabel / ilabel:=olabel;
// weights and gradient pass through.  (transducer.py:228,269,273)
template <typename R>
GraphT<R> project(const GraphT<R>& g, bool input) {
  auto gf = [](std::vector<GraphT<R>>& in, GraphT<R>& d) { in[0].addGrad(d); };
  GraphT<R> out(gf, {g});
  for (int n = 0; n < g.numNodes(); ++n) out.addNode(g.isStart(n), g.isAccept(n));
  for (int a = 0; a < g.numArcs(); ++a) {
    int l = input ? g.ilabel(a) : g.olabel(a);
    out.addArc(g.srcNode(a), g.dstNode(a), l, l, g.weight(a));
  }
  return out;
}

// remove(g, ilabel, olabel): keeps nodes that are start or have at least one
// in-arc that is not (ilabel:olabel); each kept node absorbs the closure over
// matching arcs (inheriting accept-ness) and copies every other out-arc with
// weight 0.  Not differentiable.  (transducer.py:222,228,269,274)
template <typename R>
GraphT<R> removeLabel(const GraphT<R>& g, int il, int ol) {
  auto match = [&](int a) { return g.ilabel(a) == il && g.olabel(a) == ol; };
  std::vector<int> nodes(g.numNodes(), -1);
  GraphT<R> out;
  for (int n = 0; n < g.numNodes(); ++n) {
    const auto& in = g.in(n);
    if (g.isStart(n) || !std::all_of(in.begin(), in.end(), match))
      nodes[n] = out.addNode(g.isStart(n));
  }
  std::queue<int> toExplore;
  std::set<int> reachable;
  for (int n = 0; n < g.numNodes(); ++n) {
    int curr = nodes[n];
    if (curr >= 0) {
      toExplore.push(n);
      reachable.insert(n);
    }
    while (!toExplore.empty()) {
      int next = toExplore.front();
      toExplore.pop();
      if (g.isAccept(next)) out.makeAccept(curr);
      for (int a : g.out(next)) {
        int dn = g.dstNode(a);
        if (match(a)) {
          if (!reachable.count(dn)) {
            toExplore.push(dn);
            reachable.insert(dn);
          }
        } else {
          out.addArc(curr, nodes[dn], g.ilabel(a), g.olabel(a));
        }
      }
    }
    reachable.clear();
  }
  return out;
}

// ---------------------------------------------------------------------------
// composition
// ---------------------------------------------------------------------------
namespace detail {

// Enumerates pairs (arc of g1 at node n1, arc of g2 at node n2) with
// g1.olabel == g2.ilabel, over out-arcs (or in-arcs when matchIn).
template <typename R>
struct Matcher {
  const GraphT<R>& g1;
  const GraphT<R>& g2;
  int mode;  // 0 unsorted, 1 g1 sorted only (search g1), 2 g2 sorted only, 3 both
  const std::vector<int>* query{nullptr};
  const std::vector<int>* search{nullptr};
  bool searchG1{false};
  size_t qi{0}, si{0}, sBegin{0};

  Matcher(const GraphT<R>& a, const GraphT<R>& b) : g1(a), g2(b) {
    bool s1 = a.olabelSorted(), s2 = b.ilabelSorted();
    mode = (s1 && s2) ? 3 : (s1 ? 1 : (s2 ? 2 : 0));
  }
  int qlabel(int arc) const { return searchG1 ? g2.ilabel(arc) : g1.olabel(arc); }
  int slabel(int arc) const { return searchG1 ? g1.olabel(arc) : g2.ilabel(arc); }

  size_t lowerBound(size_t from, int label) const {
    auto it = std::lower_bound(
        search->begin() + from, search->end(), label,
        [this](int arc, int val) { return slabel(arc) < val; });
    return (size_t)(it - search->begin());
  }

  void match(int n1, int n2, bool matchIn = false) {
    const auto& lv = matchIn ? g1.in(n1) : g1.out(n1);
    const auto& rv = matchIn ? g2.in(n2) : g2.out(n2);
    switch (mode) {
      case 0: searchG1 = false; break;               // nested loops, g1 outer
      case 1: searchG1 = true; break;                // query g2 in list order
      case 2: searchG1 = false; break;               // query g1 in list order
      default: searchG1 = lv.size() > rv.size();     // query = the shorter list
    }
    query = searchG1 ? &rv : &lv;
    search = searchG1 ? &lv : &rv;
    qi = 0;
    sBegin = 0;
    if (mode != 0 && !query->empty()) {
      sBegin = (mode == 3) ? lowerBound(0, qlabel((*query)[0]))
                           : lowerBound(0, qlabel((*query)[0]));
    }
    si = sBegin;
  }

  bool next(int* a1, int* a2) {
    while (qi < query->size()) {
      int qa = (*query)[qi];
      int ql = qlabel(qa);
      for (; si < search->size(); ++si) {
        int sa = (*search)[si];
        int sl = slabel(sa);
        if (sl == ql) {
          *a1 = searchG1 ? sa : qa;
          *a2 = searchG1 ? qa : sa;
          ++si;
          return true;
        }
        if (mode != 0 && ql < sl) break;  // search side sorted: no more matches
      }
      ++qi;
      if (qi < query->size()) {
        if (mode == 0) sBegin = 0;
        else if (mode == 3) sBegin = lowerBound(sBegin, qlabel((*query)[qi]));
        else sBegin = lowerBound(0, qlabel((*query)[qi]));
      }
      si = sBegin;
    }
    return false;
  }
};

}  // namespace detail

template <typename R>
GraphT<R> compose(const GraphT<R>& first, const GraphT<R>& second) {
  detail::Matcher<R> m(first, second);
  const int n1 = first.numNodes();
  auto toIndex = [n1](int a, int b) { return (size_t)a + (size_t)n1 * (size_t)b; };

  // 1. co-reachability: BFS backwards from accept pairs
  std::vector<char> reachable((size_t)n1 * (size_t)second.numNodes(), 0);
  {
    std::queue<std::pair<int, int>> q;
    for (int f : first.accept())
      for (int s : second.accept()) {
        q.emplace(f, s);
        reachable[toIndex(f, s)] = 1;
      }
    while (!q.empty()) {
      auto cur = q.front();
      q.pop();
      bool epsMatched = false;
      m.match(cur.first, cur.second, true);
      int i, j;
      while (m.next(&i, &j)) {
        epsMatched = epsMatched || (first.olabel(i) == kEps);
        int u1 = first.srcNode(i), u2 = second.srcNode(j);
        auto idx = toIndex(u1, u2);
        if (!reachable[idx]) q.emplace(u1, u2);
        reachable[idx] = 1;
      }
      if (!epsMatched) {
        for (int a : first.in(cur.first)) {
          if (first.olabel(a) != kEps) {
            if (first.olabelSorted()) break; else continue;
          }
          int u1 = first.srcNode(a);
          auto idx = toIndex(u1, cur.second);
          if (!reachable[idx]) q.emplace(u1, cur.second);
          reachable[idx] = 1;
        }
        for (int a : second.in(cur.second)) {
          if (second.ilabel(a) != kEps) {
            if (second.ilabelSorted()) break; else continue;
          }
          int u2 = second.srcNode(a);
          auto idx = toIndex(cur.first, u2);
          if (!reachable[idx]) q.emplace(cur.first, u2);
          reachable[idx] = 1;
        }
      }
    }
  }

  // 2. forward build in BFS discovery order
  auto gradInfo = std::make_shared<std::vector<std::pair<int, int>>>();
  auto gf = [gradInfo](std::vector<GraphT<R>>& in, GraphT<R>& d) {
    bool c1 = in[0].calcGrad(), c2 = in[1].calcGrad();
    std::vector<R> g1(c1 ? in[0].numArcs() : 0, R(0));
    std::vector<R> g2(c2 ? in[1].numArcs() : 0, R(0));
    for (size_t k = 0; k < gradInfo->size(); ++k) {
      R v = d.weight((int)k);
      auto& p = (*gradInfo)[k];
      if (c1 && p.first >= 0) g1[p.first] += v;
      if (c2 && p.second >= 0) g2[p.second] += v;
    }
    if (c1) in[0].addGrad(std::move(g1));
    if (c2) in[1].addGrad(std::move(g2));
  };
  GraphT<R> out(gf, {first, second});
  std::vector<int> newNodes((size_t)n1 * (size_t)second.numNodes(), -1);
  std::queue<std::pair<int, int>> q;
  for (int s1 : first.start())
    for (int s2 : second.start()) {
      auto idx = toIndex(s1, s2);
      if (reachable[idx]) {
        newNodes[idx] = out.addNode(true, first.isAccept(s1) && second.isAccept(s2));
        q.emplace(s1, s2);
      }
    }
  while (!q.empty()) {
    auto cur = q.front();
    q.pop();
    int curNode = newNodes[toIndex(cur.first, cur.second)];
    bool epsMatched = false;
    int i, j;
    m.match(cur.first, cur.second, false);
    while (m.next(&i, &j)) {
      epsMatched = epsMatched || (first.olabel(i) == kEps);
      int d1 = first.dstNode(i), d2 = second.dstNode(j);
      auto idx = toIndex(d1, d2);
      if (!reachable[idx]) continue;
      if (newNodes[idx] < 0) {
        newNodes[idx] = out.addNode(first.isStart(d1) && second.isStart(d2),
                                    first.isAccept(d1) && second.isAccept(d2));
        q.emplace(d1, d2);
      }
      out.addArc(curNode, newNodes[idx], first.ilabel(i), second.olabel(j),
                 first.weight(i) + second.weight(j));
      gradInfo->emplace_back(i, j);
    }
    if (!epsMatched) {
      for (int a : first.out(cur.first)) {
        if (first.olabel(a) != kEps) {
          if (first.olabelSorted()) break; else continue;
        }
        int d1 = first.dstNode(a);
        auto idx = toIndex(d1, cur.second);
        if (!reachable[idx]) continue;
        if (newNodes[idx] < 0) {
          newNodes[idx] = out.addNode(first.isStart(d1) && second.isStart(cur.second),
                                      first.isAccept(d1) && second.isAccept(cur.second));
          q.emplace(d1, cur.second);
        }
        out.addArc(curNode, newNodes[idx], first.ilabel(a), kEps, first.weight(a));
        gradInfo->emplace_back(a, -1);
      }
      for (int a : second.out(cur.second)) {
        if (second.ilabel(a) != kEps) {
          if (second.ilabelSorted()) break; else continue;
        }
        int d2 = second.dstNode(a);
        auto idx = toIndex(cur.first, d2);
        if (!reachable[idx]) continue;
        if (newNodes[idx] < 0) {
          newNodes[idx] = out.addNode(first.isStart(cur.first) && second.isStart(d2),
                                      first.isAccept(cur.first) && second.isAccept(d2));
          q.emplace(cur.first, d2);
        }
        out.addArc(curNode, newNodes[idx], kEps, second.olabel(a), second.weight(a));
        gradInfo->emplace_back(-1, a);
      }
    }
  }
  return out;
}

// ---------------------------------------------------------------------------
// shortest distance (forward_score / viterbi_score / viterbi_path)
// ---------------------------------------------------------------------------
namespace detail {

template <typename R>
inline R reduceScores(const std::vector<R>& in, R maxScore, bool tropical) {
  const R inf = std::numeric_limits<R>::infinity();
  if (in.empty()) return -inf;
  if (maxScore == inf || maxScore == -inf) return maxScore;
  if (tropical) return maxScore;
  R s = R(-1);
  for (R v : in) s += std::exp(v - maxScore);
  return maxScore + std::log1p(s);
}

// Kahn traversal. Fills scores[n]; for the tropical semiring also the best
// incoming arc per node (-1 = the start contribution). Returns Z.
template <typename R>
R kahnForward(const GraphT<R>& g, bool tropical, std::vector<R>& scores,
              std::vector<int>* bestArc, int* bestAccept) {
  const R inf = std::numeric_limits<R>::infinity();
  const int N = g.numNodes();
  scores.assign(N, -inf);
  if (bestArc) bestArc->assign(N, -1);
  std::vector<int> degree(N);
  std::queue<int> ready;
  for (int n = 0; n < N; ++n) {
    degree[n] = (int)g.in(n).size();
    if (degree[n] == 0) ready.push(n);
  }
  int visited = 0;
  std::vector<R> inScores;
  while (!ready.empty()) {
    int n = ready.front();
    ready.pop();
    ++visited;
    inScores.clear();
    R maxScore = -inf;
    int arg = -1;
    if (g.isStart(n)) {
      inScores.push_back(R(0));
      maxScore = R(0);
    }
    for (int a : g.in(n)) {
      R v = scores[g.srcNode(a)] + g.weight(a);
      inScores.push_back(v);
      if (v > maxScore) {
        maxScore = v;
        arg = a;
      }
    }
    scores[n] = reduceScores(inScores, maxScore, tropical);
    if (bestArc) (*bestArc)[n] = arg;
    for (int a : g.out(n)) {
      int d = g.dstNode(a);
      if (--degree[d] == 0) ready.push(d);
    }
  }
  if (visited != N)
    throw std::invalid_argument("[shortestDistance] graph has a cycle or self-loop");
  inScores.clear();
  R maxScore = -inf;
  int argAcc = -1;
  for (int n : g.accept()) {
    inScores.push_back(scores[n]);
    if (scores[n] > maxScore) {
      maxScore = scores[n];
      argAcc = n;
    }
  }
  if (bestAccept) *bestAccept = argAcc;
  return reduceScores(inScores, maxScore, tropical);
}

// reverse sweep: arc posteriors scaled by the incoming delta
template <typename R>
std::vector<R> kahnBackward(const GraphT<R>& g, bool tropical,
                            const std::vector<R>& scores,
                            const std::vector<int>& bestArc, int bestAccept,
                            R Z, R delta) {
  const R inf = std::numeric_limits<R>::infinity();
  const int N = g.numNodes();
  std::vector<R> nodeGrad(N, R(0));
  std::vector<R> arcGrad(g.numArcs(), R(0));
  if (Z == -inf || Z == inf) return arcGrad;  // infeasible: no gradient
  if (tropical) {
    int n = bestAccept;
    while (n >= 0) {
      int a = bestArc[n];
      if (a < 0) break;
      arcGrad[a] = delta;
      n = g.srcNode(a);
    }
    return arcGrad;
  }
  for (int n : g.accept()) nodeGrad[n] = delta * std::exp(scores[n] - Z);
  std::vector<int> degree(N);
  std::queue<int> ready;
  for (int n = 0; n < N; ++n) {
    degree[n] = (int)g.out(n).size();
    if (degree[n] == 0) ready.push(n);
  }
  while (!ready.empty()) {
    int n = ready.front();
    ready.pop();
    R sn = scores[n];
    for (int a : g.in(n)) {
      int u = g.srcNode(a);
      if (sn != -inf && nodeGrad[n] != R(0)) {
        R ag = nodeGrad[n] * std::exp(scores[u] + g.weight(a) - sn);
        arcGrad[a] = ag;
        nodeGrad[u] += ag;
      }
      if (--degree[u] == 0) ready.push(u);
    }
  }
  return arcGrad;
}

}  // namespace detail

template <typename R>
GraphT<R> shortestDistance(const GraphT<R>& g, bool tropical) {
  auto scores = std::make_shared<std::vector<R>>();
  auto bestArc = std::make_shared<std::vector<int>>();
  int bestAccept = -1;
  R Z = detail::kahnForward(g, tropical, *scores, bestArc.get(), &bestAccept);
  auto gf = [scores, bestArc, bestAccept, Z, tropical](std::vector<GraphT<R>>& in,
                                                       GraphT<R>& d) {
    in[0].addGrad(detail::kahnBackward(in[0], tropical, *scores, *bestArc,
                                       bestAccept, Z, d.item()));
  };
  GraphT<R> out(gf, {g});
  out.addNode(true);
  out.addNode(false, true);
  out.addArc(0, 1, 0, 0, Z);
  return out;
}

template <typename R>
GraphT<R> viterbiPath(const GraphT<R>& g) {
  std::vector<R> scores;
  std::vector<int> bestArc;
  int bestAccept = -1;
  detail::kahnForward(g, true, scores, &bestArc, &bestAccept);
  std::vector<int> arcs;
  int n = bestAccept;
  while (n >= 0) {
    int a = bestArc[n];
    if (a < 0) break;
    arcs.push_back(a);
    n = g.srcNode(a);
  }
  std::reverse(arcs.begin(), arcs.end());
  auto path = std::make_shared<std::vector<int>>(arcs);
  auto gf = [path](std::vector<GraphT<R>>& in, GraphT<R>& d) {
    std::vector<R> gr(in[0].numArcs(), R(0));
    for (size_t k = 0; k < path->size(); ++k) gr[(*path)[k]] += d.weight((int)k);
    in[0].addGrad(std::move(gr));
  };
  GraphT<R> out(gf, {g});
  if (bestAccept < 0) return out;
  out.addNode(true, arcs.empty());
  for (size_t k = 0; k < arcs.size(); ++k) {
    int a = arcs[k];
    out.addNode(false, k + 1 == arcs.size());
    out.addArc((int)k, (int)k + 1, g.ilabel(a), g.olabel(a), g.weight(a));
  }
  return out;
}

// ---------------------------------------------------------------------------
// reverse mode over the tape
// ---------------------------------------------------------------------------
template <typename R>
void backward(GraphT<R> g, const GraphT<R>& seed, bool retainGraph = false) {
  std::unordered_set<std::uintptr_t> seen;
  std::vector<GraphT<R>> tape;
  std::function<void(GraphT<R>&)> visit = [&](GraphT<R>& x) {
    if (!x.calcGrad() || seen.count(x.id())) return;
    seen.insert(x.id());
    for (auto& in : x.inputs()) visit(in);
    tape.push_back(x);
  };
  visit(g);
  g.addGrad(seed);
  for (auto it = tape.rbegin(); it != tape.rend(); ++it) {
    if (it->gradFunc()) {
      if (it->hasGrad()) it->gradFunc()(it->inputs(), it->grad());
      if (!retainGraph) {
        it->zeroGrad();
        it->inputs().clear();
        it->gradFunc() = nullptr;
      }
    }
  }
}

template <typename R>
void backward(GraphT<R> g, bool retainGraph = false) {
  GraphT<R> seed = GraphT<R>::deepCopy(g);
  std::fill(seed.weights().begin(), seed.weights().end(), R(1));
  backward(g, seed, retainGraph);
}

// ---------------------------------------------------------------------------
// comparisons
// ---------------------------------------------------------------------------
template <typename R>
bool equal(const GraphT<R>& a, const GraphT<R>& b) {
  if (a.numNodes() != b.numNodes() || a.numArcs() != b.numArcs()) return false;
  for (int n = 0; n < a.numNodes(); ++n)
    if (a.isStart(n) != b.isStart(n) || a.isAccept(n) != b.isAccept(n)) return false;
  // same arcs between the same node numbers, irrespective of insertion order
  auto key = [](const GraphT<R>& g, int k) {
    return std::make_tuple(g.srcNode(k), g.dstNode(k), g.ilabel(k), g.olabel(k), g.weight(k));
  };
  std::vector<std::tuple<int, int, int, int, R>> ka, kb;
  for (int k = 0; k < a.numArcs(); ++k) {
    ka.push_back(key(a, k));
    kb.push_back(key(b, k));
  }
  std::sort(ka.begin(), ka.end());
  std::sort(kb.begin(), kb.end());
  return ka == kb;
}

namespace detail {
template <typename R>
bool isoRec(const GraphT<R>& a, const GraphT<R>& b, int na, int nb,
            std::set<std::pair<int, int>>& visited) {
  auto key = std::make_pair(na, nb);
  if (visited.count(key)) return true;
  if (a.isStart(na) != b.isStart(nb) || a.isAccept(na) != b.isAccept(nb) ||
      a.out(na).size() != b.out(nb).size() || a.in(na).size() != b.in(nb).size())
    return false;
  visited.insert(key);
  std::vector<char> used(b.out(nb).size(), 0);
  for (int ea : a.out(na)) {
    bool ok = false;
    for (size_t k = 0; k < b.out(nb).size(); ++k) {
      if (used[k]) continue;
      int eb = b.out(nb)[k];
      if (a.ilabel(ea) != b.ilabel(eb) || a.olabel(ea) != b.olabel(eb) ||
          a.weight(ea) != b.weight(eb))
        continue;
      auto snapshot = visited;
      if (isoRec(a, b, a.dstNode(ea), b.dstNode(eb), visited)) {
        used[k] = 1;
        ok = true;
        break;
      }
      visited = snapshot;
    }
    if (!ok) {
      visited.erase(key);
      return false;
    }
  }
  return true;
}
}  // namespace detail

template <typename R>
bool isomorphic(const GraphT<R>& a, const GraphT<R>& b) {
  if (a.numNodes() != b.numNodes() || a.numArcs() != b.numArcs() ||
      a.start().size() != b.start().size() || a.accept().size() != b.accept().size())
    return false;
  if (a.numNodes() == 0) return true;
  // try to match the first start node of a with each start node of b
  if (a.start().empty()) return equal(a, b);
  for (int sb : b.start()) {
    std::set<std::pair<int, int>> visited;
    if (detail::isoRec(a, b, a.start()[0], sb, visited)) return true;
  }
  return false;
}

// ---------------------------------------------------------------------------
// text / binary io (tests/trans_backoff_test.txt: line 1 start nodes, line 2
// accept nodes, then "src dst ilabel [olabel [weight]]")
// ---------------------------------------------------------------------------
template <typename R>
GraphT<R> loadTxt(std::istream& in) {
  auto splitInts = [](const std::string& line) {
    std::vector<int> v;
    std::istringstream ss(line);
    int x;
    while (ss >> x) v.push_back(x);
    return v;
  };
  std::string line;
  if (!std::getline(in, line)) throw std::invalid_argument("[loadtxt] empty input");
  auto starts = splitInts(line);
  if (!std::getline(in, line)) throw std::invalid_argument("[loadtxt] missing accept line");
  auto accepts = splitInts(line);
  struct A { int s, d, il, ol; R w; };
  std::vector<A> arcs;
  int maxNode = -1;
  for (int s : starts) maxNode = std::max(maxNode, s);
  for (int s : accepts) maxNode = std::max(maxNode, s);
  while (std::getline(in, line)) {
    std::istringstream ss(line);
    std::vector<std::string> tok;
    std::string t;
    while (ss >> t) tok.push_back(t);
    if (tok.empty()) continue;
    if (tok.size() < 3 || tok.size() > 5) throw std::invalid_argument("[loadtxt] bad arc line");
    A a;
    a.s = std::stoi(tok[0]);
    a.d = std::stoi(tok[1]);
    a.il = std::stoi(tok[2]);
    a.ol = tok.size() > 3 ? std::stoi(tok[3]) : a.il;
    a.w = tok.size() > 4 ? (R)std::stod(tok[4]) : R(0);
    maxNode = std::max(maxNode, std::max(a.s, a.d));
    arcs.push_back(a);
  }
  std::vector<char> isS(maxNode + 1, 0), isA(maxNode + 1, 0);
  for (int s : starts) isS[s] = 1;
  for (int s : accepts) isA[s] = 1;
  GraphT<R> g;
  for (int n = 0; n <= maxNode; ++n) g.addNode(isS[n], isA[n]);
  for (auto& a : arcs) g.addArc(a.s, a.d, a.il, a.ol, a.w);
  return g;
}

template <typename R>
void saveTxt(std::ostream& out, const GraphT<R>& g) {
  auto dump = [&](const std::vector<int>& v) {
    for (size_t i = 0; i < v.size(); ++i) out << (i ? " " : "") << v[i];
    out << "\n";
  };
  dump(g.start());
  dump(g.accept());
  out.precision(9);
  for (int a = 0; a < g.numArcs(); ++a)
    out << g.srcNode(a) << " " << g.dstNode(a) << " " << g.ilabel(a) << " "
        << g.olabel(a) << " " << g.weight(a) << "\n";
}

// binary, as recalled from gtn/utils.cpp saveGraph / loadGraph (GTN is not in the
// container, so the field order is "parity unpinned"): int32 numNodes, numStart,
// numAccept; the start ids; the accept ids; int32 numArcs; then per arc
// {src, dst, ilabel, olabel} int32 and weight float32.
template <typename R>
void saveBin(std::ostream& out, const GraphT<R>& g) {
  auto wi = [&](int v) { out.write(reinterpret_cast<const char*>(&v), 4); };
  wi(g.numNodes());
  wi((int)g.start().size());
  wi((int)g.accept().size());
  for (int s : g.start()) wi(s);
  for (int s : g.accept()) wi(s);
  wi(g.numArcs());
  for (int a = 0; a < g.numArcs(); ++a) {
    wi(g.srcNode(a));
    wi(g.dstNode(a));
    wi(g.ilabel(a));
    wi(g.olabel(a));
    float w = (float)g.weight(a);
    out.write(reinterpret_cast<const char*>(&w), 4);
  }
}

template <typename R>
GraphT<R> loadBin(std::istream& in) {
  auto ri = [&]() {
    int v = 0;
    in.read(reinterpret_cast<char*>(&v), 4);
    if (!in) throw std::invalid_argument("[load] truncated graph file");
    return v;
  };
  int nn = ri(), ns = ri(), na = ri();
  if (nn < 0 || ns < 0 || na < 0 || ns > nn || na > nn) throw std::invalid_argument("[load] corrupt header");
  std::vector<char> isS(nn, 0), isA(nn, 0);
  for (int i = 0; i < ns; ++i) isS.at(ri()) = 1;
  for (int i = 0; i < na; ++i) isA.at(ri()) = 1;
  int narcs = ri();
  if (narcs < 0) throw std::invalid_argument("[load] corrupt header");
  GraphT<R> g;
  for (int n = 0; n < nn; ++n) g.addNode(isS[n], isA[n]);
  for (int a = 0; a < narcs; ++a) {
    int s = ri(), d = ri(), il = ri(), ol = ri();
    float w;
    in.read(reinterpret_cast<char*>(&w), 4);
    if (!in) throw std::invalid_argument("[load] truncated graph file");
    g.addArc(s, d, il, ol, (R)w);
  }
  return g;
}

}  // namespace og
