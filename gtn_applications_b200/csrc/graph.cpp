// Host-side WFST container and the small-graph algebra the transducer / STC criteria need
// before the GPU lattice kernels run: the per-utterance graphs of
// criterions/transducer.py:260-281 (chain o lexicon, project, epsilon removal, tokens o
// decompositions) are built here in C++, on a pool of host threads, and packed into the
// CSR batch layout of wfst_acceptor_batch_t.  This replaces the corresponding calls into
// the external GTN library (gtn.Graph / compose / intersect / remove / project_* /
// arc_sort / load / loadtxt); node and arc numbering follows GTN's construction order so
// that graph indices are bit-identical to the reference run on the oracle.
//
// Storage is structure-of-arrays (src / dst / ilabel / olabel / weight per arc) with
// per-node in/out lists kept as index vectors that arc_sort permutes; composition walks
// state pairs breadth-first from the start pairs after a backward co-reachability pass.
#include <algorithm>
#include <array>
#include <climits>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <unistd.h>
#include <functional>
#include <list>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/wfst_b200.h"

namespace wfst {
void set_error(const char* fmt, ...);

namespace {

constexpr int kEpsilon = -1;

// adjacency list of one node: up to 6 arc ids inline, heap beyond that.  The graphs built per
// utterance (compose / remove / project chains) have thousands of nodes of degree 1-5;
// std::vector per node made malloc/free the dominant cost of building them.
class AdjList {
 public:
  AdjList() : sz_(0), cap_(kInline) {}
  AdjList(const AdjList& o) : sz_(0), cap_(kInline) { assign(o); }
  AdjList(AdjList&& o) noexcept : sz_(o.sz_), cap_(o.cap_) {
    if (o.cap_ == kInline) std::memcpy(u_.inl, o.u_.inl, sizeof(u_.inl));
    else { u_.heap = o.u_.heap; o.cap_ = kInline; }
    o.sz_ = 0;
  }
  AdjList& operator=(const AdjList& o) { if (this != &o) { sz_ = 0; assign(o); } return *this; }
  AdjList& operator=(AdjList&& o) noexcept {
    if (this != &o) {
      if (cap_ != kInline) std::free(u_.heap);
      sz_ = o.sz_; cap_ = o.cap_;
      if (o.cap_ == kInline) std::memcpy(u_.inl, o.u_.inl, sizeof(u_.inl));
      else { u_.heap = o.u_.heap; o.cap_ = kInline; }
      o.sz_ = 0;
    }
    return *this;
  }
  ~AdjList() { if (cap_ != kInline) std::free(u_.heap); }
  size_t size() const { return sz_; }
  bool empty() const { return sz_ == 0; }
  int32_t* begin() { return data(); }
  int32_t* end() { return data() + sz_; }
  const int32_t* begin() const { return data(); }
  const int32_t* end() const { return data() + sz_; }
  int32_t operator[](size_t i) const { return data()[i]; }
  int32_t& operator[](size_t i) { return data()[i]; }
  void push_back(int32_t v) {
    if (sz_ == cap_) {
      const uint32_t ncap = cap_ * 2;
      int32_t* p = static_cast<int32_t*>(std::malloc(sizeof(int32_t) * ncap));
      std::memcpy(p, data(), sizeof(int32_t) * sz_);
      if (cap_ != kInline) std::free(u_.heap);
      u_.heap = p; cap_ = ncap;
    }
    data()[sz_++] = v;
  }
 private:
  static constexpr uint32_t kInline = 6;
  int32_t* data() { return cap_ == kInline ? u_.inl : u_.heap; }
  const int32_t* data() const { return cap_ == kInline ? u_.inl : u_.heap; }
  void assign(const AdjList& o) { for (int32_t v : o) push_back(v); }
  union { int32_t inl[kInline]; int32_t* heap; } u_;
  uint32_t sz_, cap_;
};

struct HostGraph {
  std::vector<uint8_t> flags;           // bit0 start, bit1 accept
  std::vector<int32_t> starts, accepts; // in insertion order
  std::vector<int32_t> src, dst, il, ol;
  std::vector<float> w;
  std::vector<AdjList> in, out;
  bool il_sorted = false, ol_sorted = false;
  bool calc_grad = false;
  // identity for the alignment-graph cache: a process-wide serial number; frozen graphs are
  // shared with that cache and must not be modified through a handle
  uint64_t serial = next_serial();
  bool frozen = false;
  static uint64_t next_serial() {
    static std::atomic<uint64_t> n{1};
    return n.fetch_add(1);
  }
  // provenance of a composed graph: arc k came from (a1[k], a2[k]) (-1 = epsilon move)
  std::vector<int32_t> prov1, prov2;
  // labels of the out-lists in list order, flattened (out_off[n] .. out_off[n+1]): built on
  // demand for big static graphs (token graphs: 10^6 arcs) so that the matcher's binary
  // searches run over contiguous memory instead of chasing arc ids
  mutable std::mutex cache_mu;
  mutable std::atomic<bool> cache_ok{false};
  mutable std::vector<int32_t> out_off, out_il_flat, out_ol_flat;
  // for nodes whose out-list is olabel-sorted over a small dense label range: first position
  // of every label (lut_off[n] < 0: no table for node n; entry = position or -1)
  // per arc {olabel, dst, ilabel, weight} in one 16-byte record (big graphs only): the composition
  // touches all four for every move it makes, from four 4 MB arrays otherwise
  struct ArcRec { int32_t ol, dst, il; float w; };
  mutable std::vector<ArcRec> arc_rec;
  mutable std::vector<int64_t> lut_off;
  mutable std::vector<int32_t> lut;
  mutable int32_t lut_min = 0, lut_range = 0;
  void ensure_label_cache() const {
    if (cache_ok.load(std::memory_order_acquire)) return;
    std::lock_guard<std::mutex> lk(cache_mu);
    if (cache_ok.load(std::memory_order_relaxed)) return;
    out_off.assign(out.size() + 1, 0);
    for (size_t n = 0; n < out.size(); ++n) out_off[n + 1] = out_off[n] + (int32_t)out[n].size();
    out_il_flat.resize(src.size());
    out_ol_flat.resize(src.size());
    for (size_t n = 0; n < out.size(); ++n) {
      int32_t k = out_off[n];
      for (int32_t a : out[n]) { out_il_flat[k] = il[a]; out_ol_flat[k] = ol[a]; ++k; }
    }
    arc_rec.resize(src.size());
    for (size_t a = 0; a < src.size(); ++a) arc_rec[a] = {ol[a], dst[a], il[a], w[a]};
    lut_off.assign(out.size(), -1);
    lut.clear();
    if (ol_sorted && !ol.empty()) {
      const auto mm = std::minmax_element(ol.begin(), ol.end());
      lut_min = *mm.first;
      lut_range = *mm.second - *mm.first + 1;
      if (lut_range <= 4096) {
        for (size_t n = 0; n < out.size(); ++n) {
          const int32_t deg = out_off[n + 1] - out_off[n];
          if (deg < 64 || (int64_t)lut.size() + lut_range > (int64_t)1 << 24) continue;
          lut_off[n] = (int64_t)lut.size();
          lut.resize(lut.size() + lut_range, -1);
          int32_t* t = lut.data() + lut_off[n];
          const int32_t* labs = out_ol_flat.data() + out_off[n];
          for (int32_t k = deg - 1; k >= 0; --k) t[labs[k] - lut_min] = k;
        }
      }
    }
    cache_ok.store(true, std::memory_order_release);
  }

  int num_nodes() const { return (int)flags.size(); }
  int num_arcs() const { return (int)src.size(); }
  int add_node(bool start, bool accept) {
    int id = num_nodes();
    flags.push_back((uint8_t)((start ? 1 : 0) | (accept ? 2 : 0)));
    in.emplace_back();
    out.emplace_back();
    if (start) starts.push_back(id);
    if (accept) accepts.push_back(id);
    return id;
  }
  int add_arc(int s, int d, int i, int o, float weight) {
    int id = num_arcs();
    src.push_back(s); dst.push_back(d); il.push_back(i); ol.push_back(o); w.push_back(weight);
    out[s].push_back(id);
    in[d].push_back(id);
    il_sorted = ol_sorted = false;
    cache_ok.store(false, std::memory_order_relaxed);
    return id;
  }
  void make_accept(int n) {
    if (!(flags[n] & 2)) { flags[n] |= 2; accepts.push_back(n); }
  }
  void arc_sort(bool by_olabel) {
    if ((by_olabel && ol_sorted) || (!by_olabel && il_sorted)) return;
    const std::vector<int32_t>& key = by_olabel ? ol : il;
    auto less = [&key](int32_t a, int32_t b) { return key[a] < key[b]; };
    for (auto& v : in) std::stable_sort(v.begin(), v.end(), less);
    for (auto& v : out) std::stable_sort(v.begin(), v.end(), less);
    cache_ok.store(false, std::memory_order_relaxed);
    il_sorted = !by_olabel;
    ol_sorted = by_olabel;
  }
};

// ---- handle table ---------------------------------------------------------------
std::mutex g_mu;
std::vector<std::shared_ptr<HostGraph>> g_graphs;   // shared: cached alignment graphs are also owned by the cache
std::vector<int32_t> g_free;

int32_t put_shared(std::shared_ptr<HostGraph> g) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (!g_free.empty()) {
    int32_t h = g_free.back();
    g_free.pop_back();
    g_graphs[h] = std::move(g);
    return h;
  }
  g_graphs.push_back(std::move(g));
  return (int32_t)g_graphs.size() - 1;
}
int32_t put(std::unique_ptr<HostGraph> g) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (!g_free.empty()) {
    int32_t h = g_free.back();
    g_free.pop_back();
    g_graphs[h] = std::move(g);
    return h;
  }
  g_graphs.push_back(std::move(g));
  return (int32_t)g_graphs.size() - 1;
}
HostGraph* get(int32_t h) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (h < 0 || h >= (int32_t)g_graphs.size() || !g_graphs[h]) return nullptr;
  return g_graphs[h].get();
}

// ---- composition ----------------------------------------------------------------
// open-addressing map (state-pair key -> dense index); std::unordered_map spends most of a
// composition in malloc/free of its nodes
struct PairIndexMap {
  std::vector<size_t> keys;
  std::vector<int32_t> vals;
  size_t mask = 0, used = 0;
  static constexpr size_t kEmpty = ~(size_t)0;
  explicit PairIndexMap(size_t cap_log2 = 16) { keys.assign((size_t)1 << cap_log2, kEmpty); vals.resize(keys.size()); mask = keys.size() - 1; }
  static size_t hash(size_t k) { return (k * 0x9E3779B97F4A7C15ull) >> 17; }
  void grow() {
    std::vector<size_t> ok; std::vector<int32_t> ov;
    ok.swap(keys); ov.swap(vals);
    keys.assign(ok.size() * 2, kEmpty); vals.resize(keys.size()); mask = keys.size() - 1;
    for (size_t i = 0; i < ok.size(); ++i)
      if (ok[i] != kEmpty) {
        size_t h = hash(ok[i]) & mask;
        while (keys[h] != kEmpty) h = (h + 1) & mask;
        keys[h] = ok[i]; vals[h] = ov[i];
      }
  }
  // returns (value, inserted)
  std::pair<int32_t, bool> emplace(size_t k, int32_t v) {
    if ((used + 1) * 2 > keys.size()) grow();
    size_t h = hash(k) & mask;
    while (keys[h] != kEmpty) {
      if (keys[h] == k) return {vals[h], false};
      h = (h + 1) & mask;
    }
    keys[h] = k; vals[h] = v; ++used;
    return {v, true};
  }
};

// Enumerates (arc of A at node na, arc of B at node nb) with A.olabel == B.ilabel in the
// order GTN's matchers produce them: nested loops when neither
// side is sorted; the unsorted side in list order with a binary search into the sorted
// side when one is; the shorter list as the query when both are.
struct PairMatcher {
  const HostGraph& A;
  const HostGraph& B;
  int mode;  // 0 none sorted, 1 A sorted (query B), 2 B sorted (query A), 3 both
  bool cache_a, cache_b;   // flat label arrays available (big graphs only)
  // per query-side node with a long unsorted out-list: (label, position) sorted — lets a
  // short sorted list on the other side find its partners without walking the long list
  mutable std::unordered_map<int, std::vector<std::pair<int32_t, int32_t>>> qindex;
  mutable std::vector<std::array<int32_t, 3>> hits;
  static constexpr size_t kBigGraph = 1 << 15;
  PairMatcher(const HostGraph& a, const HostGraph& b) : A(a), B(b) {
    mode = (a.ol_sorted && b.il_sorted) ? 3 : (a.ol_sorted ? 1 : (b.il_sorted ? 2 : 0));
    cache_a = a.src.size() >= kBigGraph;
    cache_b = b.src.size() >= kBigGraph;
    if (cache_a) a.ensure_label_cache();
    if (cache_b) b.ensure_label_cache();
  }
  template <class F>
  void for_each(int na, int nb, bool incoming, F&& emit) const {
    const AdjList& la = incoming ? A.in[na] : A.out[na];
    const AdjList& lb = incoming ? B.in[nb] : B.out[nb];
    if (la.empty() || lb.empty()) return;
    if (la.size() == 1 && lb.size() == 1) {
      // chain against chain (most state pairs of target o lexicon): one comparison, any mode
      if (A.ol[la[0]] == B.il[lb[0]]) emit(la[0], lb[0]);
      return;
    }
    bool search_a;
    switch (mode) {
      case 0: search_a = false; break;
      case 1: search_a = true; break;
      case 2: search_a = false; break;
      default: search_a = la.size() > lb.size();
    }
    const AdjList& query = search_a ? lb : la;
    const AdjList& search = search_a ? la : lb;
    // labels by position in the list: from the flat cache when there is one
    const int32_t* fa = (!incoming && cache_a) ? A.out_ol_flat.data() + A.out_off[na] : nullptr;
    const int32_t* fb = (!incoming && cache_b) ? B.out_il_flat.data() + B.out_off[nb] : nullptr;
    const int32_t* fq = search_a ? fb : fa;
    const int32_t* fs = search_a ? fa : fb;
    auto qlab = [&](size_t q) { return fq ? fq[q] : (search_a ? B.il[query[q]] : A.ol[query[q]]); };
    auto slab = [&](size_t k) { return fs ? fs[k] : (search_a ? A.ol[search[k]] : B.il[search[k]]); };
    const int32_t* lut_a = (search_a && fa && mode != 0 && A.lut_off[na] >= 0) ? A.lut.data() + A.lut_off[na] : nullptr;
    if ((mode == 1 || mode == 2) && !incoming && query.size() >= 32 && search.size() * 8 < query.size()) {
      // same pairs in the same order as the loop below: partners of every label run of the
      // sorted list, then ordered by position in the query list
      const int qnode = search_a ? nb : na;
      auto ins = qindex.try_emplace(qnode);
      auto& ix = ins.first->second;
      if (ins.second) {
        ix.reserve(query.size());
        for (size_t q = 0; q < query.size(); ++q) ix.emplace_back(qlab(q), (int32_t)q);
        std::sort(ix.begin(), ix.end());
      }
      hits.clear();
      for (size_t s0 = 0; s0 < search.size();) {
        const int32_t lab = slab(s0);
        size_t s1 = s0 + 1;
        while (s1 < search.size() && slab(s1) == lab) ++s1;
        auto it = std::lower_bound(ix.begin(), ix.end(), std::make_pair(lab, (int32_t)INT32_MIN));
        for (; it != ix.end() && it->first == lab; ++it) hits.push_back({it->second, (int32_t)s0, (int32_t)s1});
        s0 = s1;
      }
      std::sort(hits.begin(), hits.end());
      for (auto& h : hits) {
        const int32_t qa = query[h[0]];
        for (int32_t k = h[1]; k < h[2]; ++k) emit(search_a ? search[k] : qa, search_a ? qa : search[k]);
      }
      return;
    }
    size_t lo = 0;
    for (size_t q = 0; q < query.size(); ++q) {
      const int32_t qa = query[q];
      const int ql = qlab(q);
      size_t k = 0;
      if (lut_a) {
        // olabel-sorted list with a first-position table: no search at all
        const int32_t off = ql - A.lut_min;
        const int32_t pos = (off >= 0 && off < A.lut_range) ? lut_a[off] : -1;
        if (pos < 0) continue;
        k = (size_t)pos;
      } else if (mode != 0) {
        // lower bound of ql in search[lo..) by label
        size_t a = (mode == 3 ? lo : 0), b = search.size();
        while (a < b) {
          const size_t mid = (a + b) / 2;
          if (slab(mid) < ql) a = mid + 1; else b = mid;
        }
        k = a;
        if (mode == 3) lo = k;
      }
      for (; k < search.size(); ++k) {
        const int32_t sa = search[k];
        const int sl = slab(k);
        if (sl == ql) emit(search_a ? sa : qa, search_a ? qa : sa);
        else if (mode != 0 && ql < sl) break;
      }
    }
  }
};

std::unique_ptr<HostGraph> compose_graphs(const HostGraph& A, const HostGraph& B) {
  PairMatcher m(A, B);
  const HostGraph::ArcRec* recA = m.cache_a ? A.arc_rec.data() : nullptr;   // hot fields of A's arcs, one cache line
  auto a_ol = [&](int32_t i) { return recA ? recA[i].ol : A.ol[i]; };
  auto a_dst = [&](int32_t i) { return recA ? recA[i].dst : A.dst[i]; };
  auto a_il = [&](int32_t i) { return recA ? recA[i].il : A.il[i]; };
  auto a_w = [&](int32_t i) { return recA ? recA[i].w : A.w[i]; };
  const size_t NA = (size_t)A.num_nodes();
  auto key = [NA](int a, int b) { return (size_t)a + NA * (size_t)b; };
  // Which state pairs can reach an accepting pair?  GTN answers this with a backward sweep
  // over ALL pairs before the forward build; with a 1000-token graph on one side that sweep
  // visits ~10^6 pairs per utterance although only a few thousand are reachable from the start.
  // Here: explore the pairs reachable from the start pairs first (same moves, in the same order,
  // as GTN's forward build), RECORD the moves, propagate co-accessibility backwards along
  // them, then replay the recorded moves of the live pairs to build the output.  The forward
  // build only ever asks about reachable pairs, so its result (node numbering, arc order,
  // provenance) is unchanged, and the arc matching runs once instead of twice.
  // The products here have ~10^6 state pairs of which a few thousand are ever touched: pairs are
  // numbered in discovery order through a hash map, everything else is indexed by that number.
  struct Move { int32_t i, j, to; };                     // arcs of A / B (-1: the other side moves on epsilon)
  PairIndexMap idx;                                      // pair key -> index in `pairs`
  std::vector<std::pair<int, int>> pairs;
  std::vector<Move> moves;
  std::vector<int32_t> mbeg;                             // moves of pair k: mbeg[k] .. mbeg[k+1]
  pairs.reserve(1 << 15);
  moves.reserve(1 << 15);
  mbeg.reserve(1 << 15);
  std::vector<uint8_t> live;                             // by pair index: can reach an accepting pair
  auto visit = [&](int a, int b) -> int32_t {
    auto ins = idx.emplace(key(a, b), (int32_t)pairs.size());
    if (ins.second) pairs.emplace_back(a, b);
    return ins.first;
  };
  for (int sa : A.starts)
    for (int sb : B.starts) visit(sa, sb);
  const size_t nstart = pairs.size();
  // One-step lookahead before a pair is entered: a non-accepting pair with no move out of it
  // cannot reach an accepting pair, so the move into it would be dropped by the
  // co-accessibility sweep anyway.  (target o lexicon: every position tries the ~40 word pieces
  // that start with its letter and all but one or two die on the next letter — 18 k pairs
  // entered per utterance where 1 k survive.)  Only checked when both out-lists are short.
  auto dead_end = [&](int a, int b) -> bool {
    if ((A.flags[a] & 2) && (B.flags[b] & 2)) return false;
    const AdjList& oa = A.out[a];
    const AdjList& ob = B.out[b];
    if (oa.size() > 4 || ob.size() > 4) return false;
    for (int32_t i : oa) if (A.ol[i] == kEpsilon) return false;
    for (int32_t j : ob) if (B.il[j] == kEpsilon) return false;
    for (int32_t i : oa)
      for (int32_t j : ob) if (A.ol[i] == B.il[j]) return false;
    return true;
  };
  for (size_t head = 0; head < pairs.size(); ++head) {
    const int ca = pairs[head].first, cb = pairs[head].second;
    mbeg.push_back((int32_t)moves.size());
    bool eps_pair = false;
    m.for_each(ca, cb, false, [&](int32_t i, int32_t j) {
      eps_pair = eps_pair || a_ol(i) == kEpsilon;
      const int da = a_dst(i);
      if (dead_end(da, B.dst[j])) return;
      moves.push_back({i, j, visit(da, B.dst[j])});
    });
    if (eps_pair) continue;
    for (int32_t i : A.out[ca]) {
      if (a_ol(i) != kEpsilon) { if (A.ol_sorted) break; else continue; }
      moves.push_back({i, -1, visit(a_dst(i), cb)});
    }
    for (int32_t j : B.out[cb]) {
      if (B.il[j] != kEpsilon) { if (B.il_sorted) break; else continue; }
      moves.push_back({-1, j, visit(ca, B.dst[j])});
    }
  }
  mbeg.push_back((int32_t)moves.size());
  {
    // reverse adjacency (CSR) over the recorded moves, then a backward sweep from the accepting pairs
    std::vector<int32_t> rptr(pairs.size() + 1, 0), radj(moves.size());
    for (auto& e : moves) ++rptr[e.to + 1];
    for (size_t k = 0; k < pairs.size(); ++k) rptr[k + 1] += rptr[k];
    {
      std::vector<int32_t> fill(rptr.begin(), rptr.end() - 1);
      for (size_t k = 0; k < pairs.size(); ++k)
        for (int32_t e = mbeg[k]; e < mbeg[k + 1]; ++e) radj[fill[moves[e].to]++] = (int32_t)k;
    }
    std::vector<int32_t> stack;
    live.assign(pairs.size(), 0);
    for (size_t k = 0; k < pairs.size(); ++k)
      if ((A.flags[pairs[k].first] & 2) && (B.flags[pairs[k].second] & 2)) { live[k] = 1; stack.push_back((int32_t)k); }
    while (!stack.empty()) {
      const int32_t k = stack.back();
      stack.pop_back();
      for (int32_t r = rptr[k]; r < rptr[k + 1]; ++r)
        if (!live[radj[r]]) { live[radj[r]] = 1; stack.push_back(radj[r]); }
    }
  }
  // forward pass: nodes numbered in breadth-first discovery order over the live pairs
  std::vector<int32_t> id(pairs.size(), -1);             // by pair index: node number in the output
  auto out = std::make_unique<HostGraph>();
  out->calc_grad = A.calc_grad || B.calc_grad;
  std::deque<int32_t> todo;
  for (size_t k = 0; k < nstart; ++k) {
    if (!live[k]) continue;
    const int sa = pairs[k].first, sb = pairs[k].second;
    id[k] = out->add_node(true, (A.flags[sa] & 2) && (B.flags[sb] & 2));
    todo.push_back((int32_t)k);
  }
  auto node_for = [&](int32_t k) -> int {
    if (!live[k]) return -1;
    if (id[k] < 0) {
      const int a = pairs[k].first, b = pairs[k].second;
      id[k] = out->add_node((A.flags[a] & 1) && (B.flags[b] & 1), (A.flags[a] & 2) && (B.flags[b] & 2));
      todo.push_back(k);
    }
    return id[k];
  };
  while (!todo.empty()) {
    const int32_t k = todo.front();
    todo.pop_front();
    const int cur = id[k];
    for (int32_t e = mbeg[k]; e < mbeg[k + 1]; ++e) {
      const Move& mv = moves[e];
      const int d = node_for(mv.to);
      if (d < 0) continue;
      if (mv.i >= 0 && mv.j >= 0) out->add_arc(cur, d, a_il(mv.i), B.ol[mv.j], a_w(mv.i) + B.w[mv.j]);
      else if (mv.j < 0) out->add_arc(cur, d, a_il(mv.i), kEpsilon, a_w(mv.i));
      else out->add_arc(cur, d, kEpsilon, B.ol[mv.j], B.w[mv.j]);
      out->prov1.push_back(mv.i);
      out->prov2.push_back(mv.j);
    }
  }
  return out;
}

// gtn.remove(g, ilabel, olabel): nodes kept = start nodes and nodes with an in-arc that is
// not (ilabel:olabel); each kept node absorbs the closure over matching arcs; weights dropped.
std::unique_ptr<HostGraph> remove_label(const HostGraph& g, int il, int ol) {
  auto hit = [&](int32_t a) { return g.il[a] == il && g.ol[a] == ol; };
  auto out = std::make_unique<HostGraph>();
  std::vector<int32_t> map(g.num_nodes(), -1);
  for (int n = 0; n < g.num_nodes(); ++n) {
    bool keep = g.flags[n] & 1;
    if (!keep)
      for (int32_t a : g.in[n])
        if (!hit(a)) { keep = true; break; }
    if (keep) map[n] = out->add_node(g.flags[n] & 1, false);
  }
  std::vector<int32_t> stamp(g.num_nodes(), -1);
  std::deque<int32_t> todo;
  for (int n = 0; n < g.num_nodes(); ++n) {
    const int cur = map[n];
    if (cur < 0) continue;
    todo.push_back(n);
    stamp[n] = n;
    while (!todo.empty()) {
      const int32_t u = todo.front();
      todo.pop_front();
      if (g.flags[u] & 2) out->make_accept(cur);
      for (int32_t a : g.out[u]) {
        const int32_t d = g.dst[a];
        if (hit(a)) {
          if (stamp[d] != n) { stamp[d] = n; todo.push_back(d); }
        } else {
          out->add_arc(cur, map[d], g.il[a], g.ol[a], 0.f);
        }
      }
    }
  }
  return out;
}

std::unique_ptr<HostGraph> project(const HostGraph& g, bool input) {
  auto out = std::make_unique<HostGraph>();
  out->calc_grad = g.calc_grad;
  for (int n = 0; n < g.num_nodes(); ++n) out->add_node(g.flags[n] & 1, g.flags[n] & 2);
  for (int a = 0; a < g.num_arcs(); ++a) {
    const int l = input ? g.il[a] : g.ol[a];
    out->add_arc(g.src[a], g.dst[a], l, l, g.w[a]);
  }
  return out;
}

std::unique_ptr<HostGraph> chain(const int32_t* seq, int n) {
  auto g = std::make_unique<HostGraph>();
  g->add_node(true, false);
  for (int i = 0; i < n; ++i) {
    g->add_node(false, i == n - 1);
    g->add_arc(i, i + 1, seq[i], seq[i], 0.f);
  }
  return g;
}

// alignment acceptor of one utterance (criterions/transducer.py:265-276)
std::unique_ptr<HostGraph> alignment_graph(const HostGraph& tokens, const HostGraph& lexicon,
                                           const int32_t* target, int n) {
  auto tgt = chain(target, n);
  tgt->arc_sort(true);
  auto decomps = remove_label(*project(*compose_graphs(*tgt, lexicon), false), kEpsilon, kEpsilon);
  decomps->arc_sort(false);
  auto align = project(*remove_label(*compose_graphs(tokens, *decomps), kEpsilon, kEpsilon), true);
  align->arc_sort(false);
  return align;
}

// gtn.viterbi_path of an acyclic graph: best[n] = first maximising in-arc in in-list order
// (strict >).  Returns null if the graph has a cycle.
std::unique_ptr<HostGraph> best_path(const HostGraph& g) {
  const int N = g.num_nodes();
  const float NEG = -INFINITY;
  std::vector<float> score(N, NEG);
  std::vector<int32_t> best(N, -1), deg(N);
  std::deque<int32_t> ready;
  for (int n = 0; n < N; ++n) { deg[n] = (int)g.in[n].size(); if (!deg[n]) ready.push_back(n); }
  int seen = 0;
  while (!ready.empty()) {
    const int n = ready.front();
    ready.pop_front();
    ++seen;
    float m = (g.flags[n] & 1) ? 0.f : NEG;
    int arg = -1;
    for (int32_t a : g.in[n]) {
      const float v = score[g.src[a]] + g.w[a];
      if (v > m) { m = v; arg = a; }
    }
    score[n] = m;
    best[n] = arg;
    for (int32_t a : g.out[n]) if (--deg[g.dst[a]] == 0) ready.push_back(g.dst[a]);
  }
  if (seen != N) return nullptr;
  int end = -1;
  float m = NEG;
  for (int n : g.accepts) if (score[n] > m) { m = score[n]; end = n; }
  std::vector<int32_t> path;
  for (int n = end; n >= 0 && best[n] >= 0; n = g.src[best[n]]) path.push_back(best[n]);
  std::reverse(path.begin(), path.end());
  auto out = std::make_unique<HostGraph>();
  if (end >= 0) {
    out->add_node(true, path.empty());
    for (size_t k = 0; k < path.size(); ++k) {
      const int32_t a = path[k];
      out->add_node(false, k + 1 == path.size());
      out->add_arc((int)k, (int)k + 1, g.il[a], g.ol[a], g.w[a]);
    }
  }
  return out;
}

// Persistent worker pool (the counterpart of gtn.parallel_for's thread pool): three
// parallel sections per transducer step made thread creation a visible cost.  Workers are
// detached and the pool is never destroyed (no static-destruction races at interpreter exit);
// a fork()ed child (its threads are gone) builds a new pool on first use.
class WorkerPool {
 public:
  static WorkerPool& instance() {
    static std::mutex mu;
    static WorkerPool* pool = nullptr;
    static pid_t owner = 0;
    std::lock_guard<std::mutex> lk(mu);
    if (!pool || owner != getpid()) {
      pool = new WorkerPool();     // a forked child leaks the parent's (thread-less) pool
      owner = getpid();
    }
    return *pool;
  }
  int size() const { return nthreads_; }
  // runs fn(0..n-1); one section at a time
  void run(int n, const std::function<void(int)>& fn) {
    std::lock_guard<std::mutex> section(run_mu_);
    {
      std::lock_guard<std::mutex> lk(mu_);
      job_ = &fn; n_ = n; next_.store(0); pending_ = nthreads_; ++generation_;
    }
    cv_.notify_all();
    std::unique_lock<std::mutex> lk(mu_);
    done_cv_.wait(lk, [&] { return pending_ == 0; });
    job_ = nullptr;
  }
 private:
  WorkerPool() {
    unsigned hw = std::thread::hardware_concurrency();
    nthreads_ = (int)(hw ? hw : 1);
    for (int t = 0; t < nthreads_; ++t) std::thread([this] { loop(); }).detach();
  }
  void loop() {
    uint64_t seen = 0;
    for (;;) {
      const std::function<void(int)>* job;
      int n;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return generation_ != seen; });
        seen = generation_;
        job = job_; n = n_;
      }
      for (int i; (i = next_.fetch_add(1)) < n;) (*job)(i);
      {
        std::lock_guard<std::mutex> lk(mu_);
        if (--pending_ == 0) done_cv_.notify_all();
      }
    }
  }
  std::mutex run_mu_, mu_;
  std::condition_variable cv_, done_cv_;
  const std::function<void(int)>* job_ = nullptr;
  std::atomic<int> next_{0};
  int n_ = 0, pending_ = 0, nthreads_ = 1;
  uint64_t generation_ = 0;
};

template <class F>
void parallel_for(int n, F&& fn) {
  if (n <= 1 || std::thread::hardware_concurrency() <= 1) {
    for (int i = 0; i < n; ++i) fn(i);
    return;
  }
  const std::function<void(int)> job = [&](int i) { fn(i); };
  WorkerPool::instance().run(n, job);
}


// ---- epsilon folding ------------------------------------------------------------------
// Epsilon-free view of an acceptor for the time-synchronous lattice kernels (the host side of
// gtn_applications_b200/epsilon.py, which documents the construction): every path
// u --eps*--> v --a--> x becomes one arc u --a--> x tied to the original arcs along it; every
// path u --eps*--> v into an accept node gives u a final-weight term tied to the path's arcs.
struct Folded {
  std::unique_ptr<HostGraph> graph;          // folded arcs, accept = has a final path
  std::vector<int64_t> arc_prov_ptr, arc_prov;   // original arc ids summed into folded arc k
  std::vector<int64_t> fin_node, fin_prov_ptr, fin_prov;   // one entry per epsilon path into an accept node
  bool ok = true;
};

Folded fold_epsilons(const HostGraph& g) {
  Folded f;
  const int N = g.num_nodes(), A = g.num_arcs();
  std::vector<std::vector<int32_t>> eps_out(N), emit_out(N);
  for (int a = 0; a < A; ++a) (g.il[a] == kEpsilon ? eps_out : emit_out)[g.src[a]].push_back(a);
  // all epsilon paths from u: (node, arc ids), in depth-first order of the out-lists
  struct Path { int32_t node; std::vector<int32_t> arcs; };
  std::vector<std::vector<Path>> closures(N);
  std::vector<uint8_t> state(N, 0);   // 0 new, 1 open, 2 done
  std::function<bool(int)> closure = [&](int u) -> bool {
    if (state[u] == 2) return true;
    if (state[u] == 1) return false;   // epsilon cycle
    state[u] = 1;
    std::vector<Path> out;
    out.push_back(Path{u, {}});
    for (int32_t a : eps_out[u]) {
      const int v = g.dst[a];
      if (!closure(v)) return false;
      for (const Path& p : closures[v]) {
        Path q{p.node, {}};
        q.arcs.reserve(p.arcs.size() + 1);
        q.arcs.push_back(a);
        q.arcs.insert(q.arcs.end(), p.arcs.begin(), p.arcs.end());
        out.push_back(std::move(q));
      }
    }
    closures[u] = std::move(out);
    state[u] = 2;
    return true;
  };
  f.graph = std::make_unique<HostGraph>();
  f.arc_prov_ptr.push_back(0);
  f.fin_prov_ptr.push_back(0);
  std::vector<uint8_t> acc(N, 0);
  struct NewArc { int32_t src, dst, lab; };
  std::vector<NewArc> arcs;
  for (int u = 0; u < N; ++u) {
    if (!closure(u)) { f.ok = false; return f; }
    for (const Path& p : closures[u]) {
      for (int32_t a : emit_out[p.node]) {
        arcs.push_back(NewArc{u, g.dst[a], g.il[a]});
        for (int32_t e : p.arcs) f.arc_prov.push_back(e);
        f.arc_prov.push_back(a);
        f.arc_prov_ptr.push_back((int64_t)f.arc_prov.size());
      }
      if (g.flags[p.node] & 2) {
        acc[u] = 1;
        f.fin_node.push_back(u);
        for (int32_t e : p.arcs) f.fin_prov.push_back(e);
        f.fin_prov_ptr.push_back((int64_t)f.fin_prov.size());
      }
    }
  }
  for (int u = 0; u < N; ++u) f.graph->add_node((g.flags[u] & 1) != 0, acc[u] != 0);
  for (const NewArc& a : arcs) f.graph->add_arc(a.src, a.dst, a.lab, a.lab, 0.f);
  return f;
}

// a batch of folded acceptors with the index arrays that tie their arcs / final weights to ONE
// vector of original parameters (epsilon.py, FoldedBatch)
struct FoldBatch {
  std::vector<int64_t> arc_seg, arc_src, fin_seg, fin_src, fin_node;
  int64_t num_arcs = 0, num_paths = 0, num_nodes = 0;
};
std::mutex g_fold_mu;
std::vector<std::unique_ptr<FoldBatch>> g_folds;

// ---- alignment-graph cache ---------------------------------------------------------
// The alignment acceptor of an utterance depends on (token graph, lexicon graph, target) only,
// and a training run meets the same targets again every epoch (the reference's
// benchmarks/transducer_benchmark.py meets them again every iteration).  Building one costs
// milliseconds of host time (two compositions), so finished graphs are kept, frozen, in an LRU
// bounded by their total number of arcs; a hit hands out another handle to the same graph.
struct AlignCache {
  struct Entry {
    uint64_t tk, lx;
    int tk_arcs, lx_arcs;
    std::vector<int32_t> target;
    std::shared_ptr<HostGraph> graph;
  };
  std::mutex mu;
  std::list<Entry> lru;   // front = most recent
  std::unordered_multimap<uint64_t, std::list<Entry>::iterator> index;
  size_t arcs = 0, cap = default_cap();
  std::atomic<uint64_t> hits{0}, misses{0};
  static size_t default_cap() {
    const char* s = std::getenv("WFST_ALIGN_CACHE_ARCS");
    return s ? (size_t)std::strtoull(s, nullptr, 10) : (size_t)16 << 20;
  }
  static uint64_t hash(uint64_t tk, uint64_t lx, const int32_t* t, int n) {
    uint64_t h = 1469598103934665603ull ^ (tk * 0x9e3779b97f4a7c15ull) ^ (lx << 32);
    for (int i = 0; i < n; ++i) { h ^= (uint32_t)t[i]; h *= 1099511628211ull; }
    return h ^ (uint64_t)n;
  }
  std::shared_ptr<HostGraph> find(const HostGraph& tk, const HostGraph& lx, const int32_t* t, int n, uint64_t h) {
    std::lock_guard<std::mutex> lk(mu);
    auto range = index.equal_range(h);
    for (auto it = range.first; it != range.second; ++it) {
      Entry& e = *it->second;
      if (e.tk == tk.serial && e.lx == lx.serial && e.tk_arcs == tk.num_arcs() && e.lx_arcs == lx.num_arcs() &&
          (int)e.target.size() == n && std::equal(e.target.begin(), e.target.end(), t)) {
        lru.splice(lru.begin(), lru, it->second);
        return e.graph;
      }
    }
    return nullptr;
  }
  void insert(const HostGraph& tk, const HostGraph& lx, const int32_t* t, int n, uint64_t h, std::shared_ptr<HostGraph> g) {
    std::lock_guard<std::mutex> lk(mu);
    const size_t a = (size_t)g->num_arcs() + (size_t)g->num_nodes();
    if (a > cap) return;
    lru.push_front(Entry{tk.serial, lx.serial, tk.num_arcs(), lx.num_arcs(), std::vector<int32_t>(t, t + n), std::move(g)});
    index.emplace(h, lru.begin());
    arcs += a;
    while (arcs > cap && !lru.empty()) {
      auto last = std::prev(lru.end());
      const uint64_t hl = hash(last->tk, last->lx, last->target.data(), (int)last->target.size());
      auto range = index.equal_range(hl);
      for (auto it = range.first; it != range.second; ++it)
        if (it->second == last) { index.erase(it); break; }
      arcs -= (size_t)last->graph->num_arcs() + (size_t)last->graph->num_nodes();
      lru.erase(last);
    }
  }
  void clear() {
    std::lock_guard<std::mutex> lk(mu);
    index.clear();
    lru.clear();
    arcs = 0;
  }
};
AlignCache& align_cache() {
  static AlignCache* c = new AlignCache();   // never destroyed (no static-destruction order issues)
  return *c;
}

}  // namespace
}  // namespace wfst

using namespace wfst;

#define GRAPH_OR_FAIL(var, h)                                   \
  HostGraph* var = get(h);                                      \
  if (!var) {                                                   \
    set_error("invalid graph handle %d", (int)(h));             \
    return WFST_ERR_INVALID;                                    \
  }

extern "C" {

int32_t wfst_graph_create(int calc_grad) {
  auto g = std::make_unique<HostGraph>();
  g->calc_grad = calc_grad != 0;
  return put(std::move(g));
}

int wfst_graph_destroy(int32_t h) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (h < 0 || h >= (int32_t)g_graphs.size() || !g_graphs[h]) return WFST_ERR_INVALID;
  g_graphs[h].reset();
  g_free.push_back(h);
  return WFST_OK;
}

int wfst_graph_destroy_many(const int32_t* handles, int n) {
  if (n < 0 || (n > 0 && !handles)) { set_error("bad arguments"); return WFST_ERR_INVALID; }
  std::vector<std::shared_ptr<HostGraph>> doomed(n);
  {
    std::lock_guard<std::mutex> lk(g_mu);
    for (int k = 0; k < n; ++k) {
      const int32_t h = handles[k];
      if (h < 0 || h >= (int32_t)g_graphs.size() || !g_graphs[h]) continue;
      doomed[k] = std::move(g_graphs[h]);
      g_free.push_back(h);
    }
  }
  parallel_for(n, [&](int k) { doomed[k].reset(); });
  return WFST_OK;
}

#define NOT_FROZEN(g)                                                                                   \
  if ((g)->frozen) {                                                                                    \
    set_error("this graph is shared with the alignment-graph cache and cannot be modified");           \
    return WFST_ERR_INVALID;                                                                            \
  }

int wfst_graph_add_node(int32_t h, int start, int accept) {
  GRAPH_OR_FAIL(g, h);
  NOT_FROZEN(g);
  return g->add_node(start != 0, accept != 0);
}

int wfst_graph_add_arc(int32_t h, int src, int dst, int ilabel, int olabel, float weight) {
  GRAPH_OR_FAIL(g, h);
  NOT_FROZEN(g);
  if (src < 0 || dst < 0 || src >= g->num_nodes() || dst >= g->num_nodes()) {
    set_error("add_arc: node index out of range");
    return WFST_ERR_INVALID;
  }
  return g->add_arc(src, dst, ilabel, olabel, weight);
}

int wfst_graph_add_arcs(int32_t h, int n, const int32_t* src, const int32_t* dst,
                        const int32_t* ilabel, const int32_t* olabel, const float* weight) {
  GRAPH_OR_FAIL(g, h);
  NOT_FROZEN(g);
  for (int k = 0; k < n; ++k) {
    if (src[k] < 0 || dst[k] < 0 || src[k] >= g->num_nodes() || dst[k] >= g->num_nodes()) {
      set_error("add_arcs: node index out of range");
      return WFST_ERR_INVALID;
    }
    g->add_arc(src[k], dst[k], ilabel[k], olabel ? olabel[k] : ilabel[k], weight ? weight[k] : 0.f);
  }
  return WFST_OK;
}

int wfst_graph_num_nodes(int32_t h) { GRAPH_OR_FAIL(g, h); return g->num_nodes(); }
int wfst_graph_num_arcs(int32_t h) { GRAPH_OR_FAIL(g, h); return g->num_arcs(); }

int wfst_graph_arc_sort(int32_t h, int by_olabel) {
  GRAPH_OR_FAIL(g, h);
  if ((by_olabel && g->ol_sorted) || (!by_olabel && g->il_sorted)) return WFST_OK;
  NOT_FROZEN(g);
  g->arc_sort(by_olabel != 0);
  return WFST_OK;
}
int wfst_graph_mark_arc_sorted(int32_t h, int by_olabel) {
  GRAPH_OR_FAIL(g, h);
  NOT_FROZEN(g);
  if (by_olabel) g->ol_sorted = true; else g->il_sorted = true;
  return WFST_OK;
}
int wfst_graph_sorted_flags(int32_t h) { GRAPH_OR_FAIL(g, h); return (g->il_sorted ? 1 : 0) | (g->ol_sorted ? 2 : 0); }
int wfst_graph_get_calc_grad(int32_t h) { GRAPH_OR_FAIL(g, h); return g->calc_grad ? 1 : 0; }
int wfst_graph_set_calc_grad(int32_t h, int v) { GRAPH_OR_FAIL(g, h); g->calc_grad = v != 0; return WFST_OK; }

int wfst_graph_set_weights(int32_t h, const float* w) {
  GRAPH_OR_FAIL(g, h);
  NOT_FROZEN(g);
  if (g->num_arcs()) std::memcpy(g->w.data(), w, sizeof(float) * g->num_arcs());
  return WFST_OK;
}
int wfst_graph_get_weights(int32_t h, float* w) {
  GRAPH_OR_FAIL(g, h);
  if (g->num_arcs()) std::memcpy(w, g->w.data(), sizeof(float) * g->num_arcs());
  return WFST_OK;
}
// src / dst / ilabel / olabel [num_arcs] (any may be NULL)
int wfst_graph_get_arcs(int32_t h, int32_t* src, int32_t* dst, int32_t* ilabel, int32_t* olabel) {
  GRAPH_OR_FAIL(g, h);
  const size_t n = sizeof(int32_t) * g->num_arcs();
  if (n) {
    if (src) std::memcpy(src, g->src.data(), n);
    if (dst) std::memcpy(dst, g->dst.data(), n);
    if (ilabel) std::memcpy(ilabel, g->il.data(), n);
    if (olabel) std::memcpy(olabel, g->ol.data(), n);
  }
  return WFST_OK;
}
int wfst_graph_get_node_flags(int32_t h, uint8_t* flags) {
  GRAPH_OR_FAIL(g, h);
  if (g->num_nodes()) std::memcpy(flags, g->flags.data(), g->num_nodes());
  return WFST_OK;
}
// per-node arc lists flattened in node order (after arc_sort this is the sorted order)
int wfst_graph_get_arc_order(int32_t h, int incoming, int32_t* order) {
  GRAPH_OR_FAIL(g, h);
  size_t k = 0;
  for (auto& v : (incoming ? g->in : g->out))
    for (int32_t a : v) order[k++] = a;
  return WFST_OK;
}
// provenance of a composed graph; fills -1 when the graph is not a composition
int wfst_graph_get_provenance(int32_t h, int32_t* arc_first, int32_t* arc_second) {
  GRAPH_OR_FAIL(g, h);
  for (int k = 0; k < g->num_arcs(); ++k) {
    arc_first[k] = k < (int)g->prov1.size() ? g->prov1[k] : -1;
    arc_second[k] = k < (int)g->prov2.size() ? g->prov2[k] : -1;
  }
  return WFST_OK;
}

int32_t wfst_graph_compose(int32_t a, int32_t b) {
  HostGraph* ga = get(a);
  HostGraph* gb = get(b);
  if (!ga || !gb) { set_error("invalid graph handle"); return WFST_ERR_INVALID; }
  return put(compose_graphs(*ga, *gb));
}
int32_t wfst_graph_remove(int32_t h, int ilabel, int olabel) {
  HostGraph* g = get(h);
  if (!g) { set_error("invalid graph handle"); return WFST_ERR_INVALID; }
  return put(remove_label(*g, ilabel, olabel));
}
int32_t wfst_graph_project(int32_t h, int input) {
  HostGraph* g = get(h);
  if (!g) { set_error("invalid graph handle"); return WFST_ERR_INVALID; }
  return put(project(*g, input != 0));
}
int32_t wfst_graph_linear(int M, int N, int calc_grad) {
  auto g = std::make_unique<HostGraph>();
  g->calc_grad = calc_grad != 0;
  g->add_node(true, M == 0);
  for (int m = 1; m <= M; ++m) {
    g->add_node(false, m == M);
    for (int n = 0; n < N; ++n) g->add_arc(m - 1, m, n, n, 0.f);
  }
  g->il_sorted = g->ol_sorted = true;
  return put(std::move(g));
}

// text format of tests/trans_backoff_test.txt: start nodes / accept nodes / arcs
int32_t wfst_graph_loadtxt(const char* path) {
  FILE* f = std::fopen(path, "r");
  if (!f) { set_error("loadtxt: cannot open %s", path); return WFST_ERR_INVALID; }
  std::vector<std::string> lines;
  char buf[4096];
  while (std::fgets(buf, sizeof(buf), f)) lines.emplace_back(buf);
  std::fclose(f);
  if (lines.size() < 2) { set_error("loadtxt: %s is truncated", path); return WFST_ERR_INVALID; }
  auto ints = [](const std::string& s) {
    std::vector<int> v;
    const char* p = s.c_str();
    char* e;
    for (long x = std::strtol(p, &e, 10); p != e; x = std::strtol(p, &e, 10)) { v.push_back((int)x); p = e; }
    return v;
  };
  auto starts = ints(lines[0]), accepts = ints(lines[1]);
  struct Arc { int s, d, i, o; float w; };
  std::vector<Arc> arcs;
  int maxn = -1;
  for (int s : starts) maxn = std::max(maxn, s);
  for (int s : accepts) maxn = std::max(maxn, s);
  for (size_t k = 2; k < lines.size(); ++k) {
    Arc a{0, 0, 0, 0, 0.f};
    double w = 0.0;
    int n = std::sscanf(lines[k].c_str(), "%d %d %d %d %lf", &a.s, &a.d, &a.i, &a.o, &w);
    if (n <= 0) continue;
    if (n < 3) { set_error("loadtxt: bad arc line %zu", k + 1); return WFST_ERR_INVALID; }
    if (n == 3) a.o = a.i;
    a.w = (float)w;
    maxn = std::max(maxn, std::max(a.s, a.d));
    arcs.push_back(a);
  }
  std::vector<uint8_t> fl(maxn + 1, 0);
  for (int s : starts) fl[s] |= 1;
  for (int s : accepts) fl[s] |= 2;
  auto g = std::make_unique<HostGraph>();
  g->calc_grad = true;
  for (int n = 0; n <= maxn; ++n) g->add_node(fl[n] & 1, fl[n] & 2);
  for (auto& a : arcs) g->add_arc(a.s, a.d, a.i, a.o, a.w);
  return put(std::move(g));
}

int wfst_graph_savetxt(int32_t h, const char* path) {
  GRAPH_OR_FAIL(g, h);
  FILE* f = std::fopen(path, "w");
  if (!f) { set_error("savetxt: cannot open %s", path); return WFST_ERR_INVALID; }
  for (size_t k = 0; k < g->starts.size(); ++k) std::fprintf(f, k ? " %d" : "%d", g->starts[k]);
  std::fprintf(f, "\n");
  for (size_t k = 0; k < g->accepts.size(); ++k) std::fprintf(f, k ? " %d" : "%d", g->accepts[k]);
  std::fprintf(f, "\n");
  for (int a = 0; a < g->num_arcs(); ++a)
    std::fprintf(f, "%d %d %d %d %.9g\n", g->src[a], g->dst[a], g->il[a], g->ol[a], g->w[a]);
  std::fclose(f);
  return WFST_OK;
}

// binary (gtn.save / gtn.load, utils.py:261): the layout of GTN's saveGraph as recalled from
// its public source (GTN is not in the build container; INTEGRATION.md says how to check it
// against a file written by the real library): int32 num_nodes, num_start, num_accept; the
// start ids; the accept ids; int32 num_arcs; then per arc {src, dst, ilabel, olabel} int32 +
// weight float32.  The loader also accepts the round-1 layout of this library (num_arcs as
// the fourth header word): both have the same length, the position of num_arcs tells them apart.
int wfst_graph_save(int32_t h, const char* path) {
  GRAPH_OR_FAIL(g, h);
  if (!path) { set_error("save: null path"); return WFST_ERR_INVALID; }
  FILE* f = std::fopen(path, "wb");
  if (!f) { set_error("save: cannot open %s", path); return WFST_ERR_INVALID; }
  int32_t hdr[3] = {g->num_nodes(), (int32_t)g->starts.size(), (int32_t)g->accepts.size()};
  int32_t narcs = g->num_arcs();
  bool ok = std::fwrite(hdr, 4, 3, f) == 3;
  ok = ok && std::fwrite(g->starts.data(), 4, g->starts.size(), f) == g->starts.size();
  ok = ok && std::fwrite(g->accepts.data(), 4, g->accepts.size(), f) == g->accepts.size();
  ok = ok && std::fwrite(&narcs, 4, 1, f) == 1;
  for (int a = 0; ok && a < narcs; ++a) {
    int32_t rec[4] = {g->src[a], g->dst[a], g->il[a], g->ol[a]};
    ok = std::fwrite(rec, 4, 4, f) == 4 && std::fwrite(&g->w[a], 4, 1, f) == 1;
  }
  ok = (std::fclose(f) == 0) && ok;
  if (!ok) { set_error("save: write to %s failed", path); return WFST_ERR_INVALID; }
  return WFST_OK;
}

int32_t wfst_graph_load(const char* path) {
  if (!path) { set_error("load: null path"); return WFST_ERR_INVALID; }
  FILE* f = std::fopen(path, "rb");
  if (!f) { set_error("load: cannot open %s", path); return WFST_ERR_INVALID; }
  std::vector<int32_t> words;
  {
    int32_t buf[4096];
    size_t n;
    while ((n = std::fread(buf, 4, 4096, f)) > 0) words.insert(words.end(), buf, buf + n);
  }
  std::fclose(f);
  auto fail = [&]() { set_error("load: %s is truncated or corrupt", path); return (int32_t)WFST_ERR_INVALID; };
  const size_t nw = words.size();
  if (nw < 4) return fail();
  const int64_t nn = words[0], ns = words[1], na = words[2];
  if (nn < 0 || ns < 0 || na < 0 || ns > nn || na > nn) return fail();
  // both layouts: 4 header words in total + ns + na + 5 * narcs
  if ((int64_t)nw < 4 + ns + na || ((int64_t)nw - 4 - ns - na) % 5 != 0) return fail();
  const int64_t narcs = ((int64_t)nw - 4 - ns - na) / 5;
  size_t ids;            // first start id
  if (words[3 + ns + na] == narcs) ids = 3;          // GTN layout
  else if (words[3] == narcs) ids = 4;               // round-1 layout of this library
  else return fail();
  const size_t arcs0 = 4 + (size_t)ns + (size_t)na;
  try {
    std::vector<uint8_t> fl((size_t)nn, 0);
    for (int64_t k = 0; k < ns; ++k) {
      const int32_t s = words[ids + k];
      if (s < 0 || s >= nn) return fail();
      fl[s] |= 1;
    }
    for (int64_t k = 0; k < na; ++k) {
      const int32_t s = words[ids + ns + k];
      if (s < 0 || s >= nn) return fail();
      fl[s] |= 2;
    }
    auto g = std::make_unique<HostGraph>();
    g->calc_grad = true;
    for (int64_t n = 0; n < nn; ++n) g->add_node(fl[n] & 1, fl[n] & 2);
    for (int64_t a = 0; a < narcs; ++a) {
      const int32_t* rec = &words[arcs0 + 5 * (size_t)a];
      if (rec[0] < 0 || rec[1] < 0 || rec[0] >= nn || rec[1] >= nn) return fail();
      float w;
      std::memcpy(&w, rec + 4, 4);
      g->add_arc(rec[0], rec[1], rec[2], rec[3], w);
    }
    return put(std::move(g));
  } catch (const std::exception& e) {   // bad_alloc on an absurd header: never across the C boundary
    set_error("load: %s: %s", path, e.what());
    return WFST_ERR_INVALID;
  }
}

// Kahn-order tropical shortest path with back-pointers; ties keep the first maximum in
// the node's in-list order (strict >).  Returns a chain graph of the best path's arcs.
int32_t wfst_graph_viterbi_path(int32_t h) {
  HostGraph* g = get(h);
  if (!g) { set_error("invalid graph handle"); return WFST_ERR_INVALID; }
  auto out = best_path(*g);
  if (!out) { set_error("viterbi_path: graph has a cycle"); return WFST_ERR_INVALID; }
  return put(std::move(out));
}

// Shortest distance of an acyclic graph from its start to its accept nodes, and its gradient
// with respect to the arc weights: gtn.forward_score (log semiring; the gradient is the arc
// posterior) or gtn.viterbi_score (tropical != 0; the gradient marks the arcs of the best path,
// ties as in viterbi_path), followed by gtn.backward of the scalar.  Host graphs only: this is
// the graph-building API's scoring (tests, small graphs), not the batched GPU path.
// arc_grad may be null; otherwise num_arcs floats.
int wfst_graph_score(int32_t h, int tropical, float* score, float* arc_grad) {
  GRAPH_OR_FAIL(g, h);
  if (!score) { set_error("null score pointer"); return WFST_ERR_INVALID; }
  const int N = g->num_nodes(), A = g->num_arcs();
  const double NEG = -INFINITY;
  std::vector<double> alpha(N, NEG);
  std::vector<int32_t> deg(N), order, best(N, -1);
  order.reserve(N);
  for (int n = 0; n < N; ++n) { deg[n] = (int)g->in[n].size(); if (!deg[n]) order.push_back(n); }
  auto lse = [](double a, double b) {
    if (a == -INFINITY) return b;
    if (b == -INFINITY) return a;
    const double m = a > b ? a : b;
    return m + std::log(std::exp(a - m) + std::exp(b - m));
  };
  for (size_t k = 0; k < order.size(); ++k) {
    const int n = order[k];
    double m = (g->flags[n] & 1) ? 0.0 : NEG;
    if (tropical) {
      for (int32_t a : g->in[n]) {
        const double v = alpha[g->src[a]] + (double)g->w[a];
        if (v > m) { m = v; best[n] = a; }
      }
    } else {
      for (int32_t a : g->in[n]) m = lse(m, alpha[g->src[a]] + (double)g->w[a]);
    }
    alpha[n] = m;
    for (int32_t a : g->out[n]) if (--deg[g->dst[a]] == 0) order.push_back(g->dst[a]);
  }
  if ((int)order.size() != N) { set_error("score: graph has a cycle"); return WFST_ERR_INVALID; }
  double z = NEG;
  int end = -1;
  for (int n : g->accepts) {
    if (tropical) { if (alpha[n] > z) { z = alpha[n]; end = n; } }
    else z = lse(z, alpha[n]);
  }
  *score = (float)z;
  if (!arc_grad) return WFST_OK;
  for (int a = 0; a < A; ++a) arc_grad[a] = 0.f;
  if (z == NEG) return WFST_OK;
  if (tropical) {
    for (int n = end; n >= 0 && best[n] >= 0; n = g->src[best[n]]) arc_grad[best[n]] = 1.f;
    return WFST_OK;
  }
  std::vector<double> beta(N, NEG);
  for (int k = N - 1; k >= 0; --k) {
    const int n = order[k];
    double m = (g->flags[n] & 2) ? 0.0 : NEG;
    for (int32_t a : g->out[n]) m = lse(m, (double)g->w[a] + beta[g->dst[a]]);
    beta[n] = m;
  }
  for (int a = 0; a < A; ++a) {
    const double v = alpha[g->src[a]] + (double)g->w[a] + beta[g->dst[a]] - z;
    arc_grad[a] = v == NEG ? 0.f : (float)std::exp(v);
  }
  return WFST_OK;
}

// Alignment -> token sequence for a batch of best alignments (Transducer.viterbi,
// criterions/transducer.py:223-233): per utterance a chain of the T frame labels is composed
// with the token graph, the best path taken, projected on its output labels and epsilons
// removed.  labels [B, T]; out [B, T] receives the tokens of utterance b at out[b*T ..],
// out_counts [B] their number.  Runs on host threads (the reference does this per utterance in
// Python).  The token graph must be ilabel-sorted by the caller (transducer.py:222).
int wfst_transducer_decode_paths(int32_t tokens, const int32_t* labels, int B, int T,
                                 int32_t* out, int32_t* out_counts) {
  HostGraph* tk = get(tokens);
  if (!tk) { set_error("invalid graph handle"); return WFST_ERR_INVALID; }
  if (B < 0 || T < 0 || (B > 0 && T > 0 && (!labels || !out)) || (B > 0 && !out_counts)) {
    set_error("bad arguments");
    return WFST_ERR_INVALID;
  }
  std::atomic<int> bad{0};
  parallel_for(B, [&](int b) {
    auto ch = chain(labels + (size_t)b * T, T);
    if (T == 0) ch->make_accept(0);
    auto best = best_path(*compose_graphs(*ch, *tk));
    if (!best) { bad = 1; out_counts[b] = 0; return; }
    auto res = remove_label(*project(*best, false), kEpsilon, kEpsilon);
    const int n = std::min(res->num_arcs(), T);
    for (int k = 0; k < n; ++k) out[(size_t)b * T + k] = res->il[k];
    out_counts[b] = n;
  });
  if (bad) { set_error("decode: composed graph has a cycle"); return WFST_ERR_INVALID; }
  return WFST_OK;
}

// Builds the alignment acceptor of every utterance of a batch on host threads
// (criterions/transducer.py:260-276; the reference runs this under gtn.parallel_for
// with the GIL taken for every Python-side step).  out_handles [B].
int wfst_transducer_alignment_graphs(int32_t tokens, int32_t lexicon, const int32_t* targets,
                                     const int32_t* target_offsets, int B, int32_t* out_handles) {
  HostGraph* tk = get(tokens);
  HostGraph* lx = get(lexicon);
  if (!tk || !lx) { set_error("invalid graph handle"); return WFST_ERR_INVALID; }
  tk->arc_sort(true);  // Transducer.forward: self.tokens.arc_sort(True) (transducer.py:188)
  std::vector<std::shared_ptr<HostGraph>> built(B);
  AlignCache& cache = align_cache();
  const bool use_cache = cache.cap > 0;
  parallel_for(B, [&](int b) {
    const int32_t* t = targets + target_offsets[b];
    const int n = target_offsets[b + 1] - target_offsets[b];
    uint64_t h = 0;
    if (use_cache) {
      h = AlignCache::hash(tk->serial, lx->serial, t, n);
      built[b] = cache.find(*tk, *lx, t, n, h);
      if (built[b]) { cache.hits.fetch_add(1, std::memory_order_relaxed); return; }
      cache.misses.fetch_add(1, std::memory_order_relaxed);
    }
    std::shared_ptr<HostGraph> g = alignment_graph(*tk, *lx, t, n);
    if (use_cache) {
      g->frozen = true;
      cache.insert(*tk, *lx, t, n, h, g);
    }
    built[b] = std::move(g);
  });
  for (int b = 0; b < B; ++b) out_handles[b] = put_shared(std::move(built[b]));
  return WFST_OK;
}

// STC acceptors of a whole batch (criterions/stc.py:22-64, create_stc_graph), built on host
// threads with the node and arc order of the reference's builder, then ilabel-sorted: a CTC-like
// chain (self loops on blank states only, unconditional skips) plus one <star> node per gap whose
// entering / looping arcs carry log(prob).  Labels must lie in [0, star_idx).  out_handles [B].
int wfst_stc_graphs(const int32_t* targets, const int32_t* target_offsets, int B, int star_idx,
                    float log_prob, int blank_idx, int32_t* out_handles) {
  if (!target_offsets || !out_handles || B < 0 || (!targets && B > 0 && target_offsets[B] > 0)) {
    set_error("stc graphs: null pointer argument");
    return WFST_ERR_INVALID;
  }
  for (int b = 0; b < B; ++b)
    for (int k = target_offsets[b]; k < target_offsets[b + 1]; ++k)
      if (targets[k] < 0 || targets[k] >= star_idx) {
        set_error("target label outside [0, %d)", star_idx);
        return WFST_ERR_INVALID;
      }
  std::vector<std::unique_ptr<HostGraph>> built(B);
  parallel_for(B, [&](int b) {
    const int32_t* y = targets + target_offsets[b];
    const int L = target_offsets[b + 1] - target_offsets[b];
    auto g = std::make_unique<HostGraph>();
    g->calc_grad = false;
    const int n_states = 2 * L + 1;
    for (int s = 0; s < n_states; ++s) {
      g->add_node(s == 0, s >= n_states - 2);
      const int lab = (s % 2) ? y[(s - 1) / 2] : blank_idx;
      if (lab == blank_idx) g->add_arc(s, s, lab, lab, 0.f);
      if (s > 0) g->add_arc(s - 1, s, lab, lab, 0.f);
      if ((s % 2) && s > 1) g->add_arc(s - 2, s, lab, lab, 0.f);
    }
    for (int k = 0; k <= L; ++k) {
      const int prev_tok = 2 * k - 1, prev_blank = 2 * k;
      const int c = g->add_node(false, k == L);
      const int idx = (k == L) ? star_idx : star_idx + y[k];
      if (prev_tok >= 0) g->add_arc(prev_tok, c, idx, idx, log_prob);
      g->add_arc(prev_blank, c, idx, idx, log_prob);
      g->add_arc(c, c, idx, idx, log_prob);
      if (k < L) g->add_arc(c, 2 * k + 1, y[k], y[k], 0.f);
      g->add_arc(c, prev_blank, blank_idx, blank_idx, 0.f);
    }
    g->arc_sort(false);
    built[b] = std::move(g);
  });
  for (int b = 0; b < B; ++b) out_handles[b] = put(std::move(built[b]));
  return WFST_OK;
}

// Capacity of the alignment-graph cache in arcs + nodes (0 disables it, < 0 keeps the current
// value); always empties the cache.  hits / misses (may be null) receive the counters since the
// last call.
int wfst_transducer_alignment_cache(long long capacity, unsigned long long* hits, unsigned long long* misses) {
  AlignCache& c = align_cache();
  if (hits) *hits = c.hits.exchange(0);
  if (misses) *misses = c.misses.exchange(0);
  c.clear();
  if (capacity >= 0) {
    std::lock_guard<std::mutex> lk(c.mu);
    c.cap = (size_t)capacity;
  }
  return WFST_OK;
}

// Transducer with an epsilon transition graph (criterions/transducer.py:279-281 with ngram > 1 or
// a loaded back-off graph): for every utterance  intersect(transitions, aligns[b])  (epsilons in
// place, as the reference), arc-sorted, then folded.  out_handles [B] receive the folded graphs
// (pack them with wfst_graph_pack); the returned fold object holds the index arrays that tie
// folded arcs and final weights to the TRANSITION graph's arcs (through the composition's
// provenance).  aligns == NULL: fold the transition graph itself (B must be 1).
int32_t wfst_fold_transitions_batch(int32_t transitions, const int32_t* aligns, int B, int32_t* out_handles) {
  HostGraph* tr = get(transitions);
  if (!tr || B < 1 || !out_handles) { set_error("bad arguments"); return WFST_ERR_INVALID; }
  std::vector<HostGraph*> al(B, nullptr);
  if (aligns) {
    for (int b = 0; b < B; ++b) {
      al[b] = get(aligns[b]);
      if (!al[b]) { set_error("invalid graph handle %d", (int)aligns[b]); return WFST_ERR_INVALID; }
    }
  } else if (B != 1) {
    set_error("folding the transition graph itself needs B == 1");
    return WFST_ERR_INVALID;
  }
  std::vector<Folded> folded(B);
  std::vector<std::vector<int32_t>> remap(B);
  parallel_for(B, [&](int b) {
    if (al[b]) {
      auto c = compose_graphs(*tr, *al[b]);
      c->arc_sort(false);
      folded[b] = fold_epsilons(*c);
      remap[b] = c->prov1;      // composed arc -> transition arc
    } else {
      folded[b] = fold_epsilons(*tr);
    }
  });
  for (int b = 0; b < B; ++b)
    if (!folded[b].ok) { set_error("epsilon cycle in acceptor"); return WFST_ERR_INVALID; }
  auto fb = std::make_unique<FoldBatch>();
  int64_t a0 = 0, p0 = 0, n0 = 0;
  for (int b = 0; b < B; ++b) {
    const Folded& f = folded[b];
    const std::vector<int32_t>& m = remap[b];
    auto orig = [&](int64_t x) { return m.empty() ? x : (int64_t)m[(size_t)x]; };
    const int64_t na = (int64_t)f.arc_prov_ptr.size() - 1, np = (int64_t)f.fin_prov_ptr.size() - 1;
    for (int64_t k = 0; k < na; ++k)
      for (int64_t q = f.arc_prov_ptr[k]; q < f.arc_prov_ptr[k + 1]; ++q) {
        fb->arc_seg.push_back(a0 + k);
        fb->arc_src.push_back(orig(f.arc_prov[q]));
      }
    for (int64_t k = 0; k < np; ++k) {
      fb->fin_node.push_back(n0 + f.fin_node[k]);
      for (int64_t q = f.fin_prov_ptr[k]; q < f.fin_prov_ptr[k + 1]; ++q) {
        fb->fin_seg.push_back(p0 + k);
        fb->fin_src.push_back(orig(f.fin_prov[q]));
      }
    }
    a0 += na; p0 += np; n0 += f.graph->num_nodes();
  }
  fb->num_arcs = a0; fb->num_paths = p0; fb->num_nodes = n0;
  for (int b = 0; b < B; ++b) out_handles[b] = put(std::move(folded[b].graph));
  std::lock_guard<std::mutex> lk(g_fold_mu);
  for (size_t i = 0; i < g_folds.size(); ++i)
    if (!g_folds[i]) { g_folds[i] = std::move(fb); return (int32_t)i; }
  g_folds.push_back(std::move(fb));
  return (int32_t)g_folds.size() - 1;
}

static FoldBatch* get_fold(int32_t h) {
  std::lock_guard<std::mutex> lk(g_fold_mu);
  if (h < 0 || h >= (int32_t)g_folds.size() || !g_folds[h]) return nullptr;
  return g_folds[h].get();
}

// sizes: {folded arcs, final paths, nodes, arc tie entries, final tie entries}
int wfst_fold_sizes(int32_t fold, int64_t* sizes) {
  FoldBatch* f = get_fold(fold);
  if (!f || !sizes) { set_error("invalid fold handle"); return WFST_ERR_INVALID; }
  sizes[0] = f->num_arcs; sizes[1] = f->num_paths; sizes[2] = f->num_nodes;
  sizes[3] = (int64_t)f->arc_seg.size(); sizes[4] = (int64_t)f->fin_seg.size();
  return WFST_OK;
}

// copies the tie arrays (int64): arc_seg / arc_src [sizes[3]], fin_seg / fin_src [sizes[4]], fin_node [sizes[1]]
int wfst_fold_fill(int32_t fold, int64_t* arc_seg, int64_t* arc_src, int64_t* fin_seg, int64_t* fin_src, int64_t* fin_node) {
  FoldBatch* f = get_fold(fold);
  if (!f) { set_error("invalid fold handle"); return WFST_ERR_INVALID; }
  auto cp = [](int64_t* dst, const std::vector<int64_t>& v) { if (dst && !v.empty()) std::memcpy(dst, v.data(), sizeof(int64_t) * v.size()); };
  cp(arc_seg, f->arc_seg); cp(arc_src, f->arc_src); cp(fin_seg, f->fin_seg); cp(fin_src, f->fin_src); cp(fin_node, f->fin_node);
  return WFST_OK;
}

int wfst_fold_destroy(int32_t fold) {
  std::lock_guard<std::mutex> lk(g_fold_mu);
  if (fold < 0 || fold >= (int32_t)g_folds.size() || !g_folds[fold]) return WFST_ERR_INVALID;
  g_folds[fold].reset();
  return WFST_OK;
}

// Sizes and packing of a batch of acceptors into the lattice kernel's layout
// (wfst_acceptor_batch_t); host arrays are filled, the caller uploads them.
int wfst_graph_pack_sizes(const int32_t* handles, int B, int32_t* total_nodes, int32_t* total_arcs,
                          int32_t* max_nodes, int32_t* max_arcs, int32_t* has_epsilon) {
  int tn = 0, ta = 0, mn = 0, ma = 0, eps = 0;
  for (int b = 0; b < B; ++b) {
    GRAPH_OR_FAIL(g, handles[b]);
    tn += g->num_nodes(); ta += g->num_arcs();
    mn = std::max(mn, g->num_nodes()); ma = std::max(ma, g->num_arcs());
    for (int a = 0; a < g->num_arcs(); ++a) eps |= (g->il[a] == kEpsilon);
  }
  *total_nodes = tn; *total_arcs = ta; *max_nodes = mn; *max_arcs = ma; *has_epsilon = eps;
  return WFST_OK;
}

int wfst_graph_pack(const int32_t* handles, int B, int32_t* node_offsets, int32_t* arc_offsets,
                    uint8_t* node_flags, int32_t* in_ptr, int32_t* in_src, int32_t* in_label,
                    int32_t* in_arc, int32_t* out_ptr, int32_t* out_dst, int32_t* out_label,
                    int32_t* out_arc, float* weights) {
  std::vector<HostGraph*> gs(B);
  node_offsets[0] = arc_offsets[0] = 0;
  for (int b = 0; b < B; ++b) {
    gs[b] = get(handles[b]);
    if (!gs[b]) { set_error("invalid graph handle"); return WFST_ERR_INVALID; }
    node_offsets[b + 1] = node_offsets[b] + gs[b]->num_nodes();
    arc_offsets[b + 1] = arc_offsets[b] + gs[b]->num_arcs();
  }
  parallel_for(B, [&](int b) {
    const HostGraph& g = *gs[b];
    const int n0 = node_offsets[b], a0 = arc_offsets[b];
    std::memcpy(node_flags + n0, g.flags.data(), g.num_nodes());
    if (g.num_arcs()) std::memcpy(weights + a0, g.w.data(), sizeof(float) * g.num_arcs());
    int32_t* ip = in_ptr + n0 + b;
    int32_t* op = out_ptr + n0 + b;
    int ki = 0, ko = 0;
    for (int v = 0; v < g.num_nodes(); ++v) {
      ip[v] = ki;
      for (int32_t a : g.in[v]) {
        in_src[a0 + ki] = g.src[a]; in_label[a0 + ki] = g.il[a]; in_arc[a0 + ki] = a; ++ki;
      }
      op[v] = ko;
      for (int32_t a : g.out[v]) {
        out_dst[a0 + ko] = g.dst[a]; out_label[a0 + ko] = g.il[a]; out_arc[a0 + ko] = a; ++ko;
      }
    }
    ip[g.num_nodes()] = ki;
    op[g.num_nodes()] = ko;
  });
  return WFST_OK;
}

}  // extern "C"
