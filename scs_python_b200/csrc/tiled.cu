// tiled.cu -- builds the tiled SpMV format of tiled.cuh from the device CSR copies (setup path,
// once per workspace, after equilibration has put the final values into the CSR arrays).
//
//   1. one pass over the rows (warp per row): segment id of every stored entry, and per segment
//      (row bin, warp, column bin) the entry count and the longest run of one row;
//   2. groups per segment = max(ceil(count / 32), longest row run); a column bin that is used by
//      any warp of a row bin gets at least one (padding) group in every warp, so that all warps
//      of the CTA step through the same sequence of x-slices;
//   3. stable radix sort of ((segment id, row mod 16) -> entry) with CUB: inside a segment the entries
//      are ordered by the bank pair of their accumulator, then row-major; entry number k of a segment
//      goes to group k mod G, lane k div G -- equal rows are consecutive and at most G long, so the rows
//      inside a group are distinct, and every bank pair is spread evenly over the groups;
//   4. host: split tall row bins into column pieces, assign the work items to the CTAs (longest
//      first onto the least loaded CTA), upload the schedule.
#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <queue>
#include <vector>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "tiled.cuh"

namespace b200 {

int tiled_env_mode() {
  const char *e = getenv("SCS_B200_TILED");
  if (!e || !*e) return -1;
  return atoi(e) != 0 ? 1 : 0;
}

namespace {

// warp per row
__global__ void __launch_bounds__(kThreads)
k_tl_count(CsrDev M, int cboff, long long koff, int ncb, unsigned *__restrict__ key, int *__restrict__ rowof,
           int *__restrict__ cnt, int *__restrict__ maxrow) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (long long rr = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); rr < M.nrows; rr += (long long)gridDim.x * wpb) {
    const int row = (int)rr;
    const int start = M.ptr[row], end = M.ptr[row + 1];
    const int rb = row / kTR, w = (row % kTR) / kTRW;
    const long long segb = ((long long)rb * kTW + w) * ncb + cboff;
    for (int k = start + lane; k < end; k += 32) {
      const int cb = M.idx[k] / kTC;
      // sort key: segment, then the accumulator's bank pair (row mod 16): dealing a segment's entries
      // round-robin over its groups then spreads every bank pair evenly over the groups, which cuts the
      // shared-memory bank conflicts of the y read-modify-write (rows of one bank class stay row-major)
      key[koff + k] = (unsigned)((segb + cb) * 16 + (row & 15));
      rowof[koff + k] = row;
      const bool runstart = (k == start) || (M.idx[k - 1] / kTC != cb);
      if (runstart) {  // length of this row's run inside the column bin: first k2 with idx >= (cb+1) C
        int lo = k + 1, hi = end;
        const long long lim = ((long long)cb + 1) * kTC;
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if ((long long)M.idx[mid] < lim) lo = mid + 1; else hi = mid;
        }
        const int len = lo - k;
        atomicAdd(cnt + segb + cb, len);
        atomicMax(maxrow + segb + cb, len);
      }
    }
  }
}

// thread per (row bin, column bin): groups of every warp's segment, and their sum
__global__ void __launch_bounds__(kThreads)
k_tl_groups(const int *__restrict__ cnt, const int *__restrict__ maxrow, int nrb, int ncb, int *__restrict__ ng,
            int *__restrict__ gcnt) {
  const long long tot = (long long)nrb * ncb;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < tot; t += (long long)gridDim.x * blockDim.x) {
    const int rb = (int)(t / ncb), cb = (int)(t % ncb);
    bool any = false;
    for (int w = 0; w < kTW; ++w) any |= cnt[((long long)rb * kTW + w) * ncb + cb] > 0;
    int sum = 0;
    for (int w = 0; w < kTW; ++w) {
      const long long seg = ((long long)rb * kTW + w) * ncb + cb;
      const int c = cnt[seg];
      int g = 0;
      if (c > 0) { g = (c + 31) >> 5; const int mr = maxrow[seg]; g = g > mr ? g : mr; }
      else if (any) g = 1;
      ng[seg] = g;
      sum += g;
    }
    gcnt[t] = sum;
  }
}

__global__ void __launch_bounds__(kThreads) k_tl_pad(unsigned *__restrict__ pk, double *__restrict__ val, long long slots) {
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < slots; t += (long long)gridDim.x * blockDim.x) {
    pk[t] = (unsigned)(kTR + (int)(t & 31)) << kTRowShift;  // private dummy accumulator of the lane, column 0
    val[t] = 0.0;
  }
}

// Where every entry of a segment goes: dstpos[p] = group * 32 + lane inside the segment's slot range.
//
// A 64-bit shared-memory access is served one half-warp at a time, and the kernel touches shared memory three
// times per entry: it reads x_s[col], reads y_s[row] and writes y_s[row].  Two lanes of a half-warp collide when
// their rows (or columns) fall into the same bank pair, i.e. agree modulo 16.  So a group of 32 slots is two
// half-slots of 16 lanes, and the entries of a segment are dealt greedily, in the order of the sort (bank pair of
// the row, then row-major), into the half-slot where they collide least: a row collision costs two extra
// wavefronts (load + store), a column collision one.  Entries of one row are consecutive in that order and must
// land in different GROUPS (the update of a group is one unordered read-modify-write): a bit mask of the groups
// the current row already uses enforces it.  If the greedy pass runs out of feasible half-slots (nearly full
// segments), or the segment has more than 32 groups, the segment falls back to the plain dealing (entry k ->
// group k mod G, chunk k div G, chunks alternating between the half-warps), which is feasible by construction.
// One thread per segment; a segment holds a few hundred entries.
__global__ void __launch_bounds__(128)
k_tl_deal(const unsigned *__restrict__ sval, long long nnz1, const int *__restrict__ sstart, const int *__restrict__ gbase,
          long long nseg, const int *__restrict__ rowof, CsrDev M1, CsrDev M2, int *__restrict__ dstpos, int greedy) {
  for (long long seg = blockIdx.x * (long long)blockDim.x + threadIdx.x; seg < nseg; seg += (long long)gridDim.x * blockDim.x) {
    const int s0 = sstart[seg], cnt = sstart[seg + 1] - s0;
    if (cnt <= 0) continue;
    const int G = gbase[seg + 1] - gbase[seg];
    bool ok = greedy && G <= 32;
    if (ok) {
      unsigned short rmask[64], cmask[64];
      unsigned char fill[64];
      const int nh = 2 * G;
      for (int h = 0; h < nh; ++h) { rmask[h] = 0; cmask[h] = 0; fill[h] = 0; }
      int prev_row = -1;
      unsigned used = 0u;
      for (int k = 0; k < cnt && ok; ++k) {
        const unsigned src = sval[s0 + k];
        const int row = rowof[src];
        const int col = (long long)src < nnz1 ? M1.idx[src] : M2.idx[src - nnz1];
        const int rc = row & 15, cc = col & 15;
        if (row != prev_row) { used = 0u; prev_row = row; }
        int best = -1, bc = 99;
        int h = k % nh;
        for (int t = 0; t < nh; ++t, h = (h + 1 == nh) ? 0 : h + 1) {
          if (fill[h] >= 16 || ((used >> (h >> 1)) & 1u)) continue;
          const int c = (((rmask[h] >> rc) & 1) << 1) + ((cmask[h] >> cc) & 1);
          if (c < bc) { bc = c; best = h; if (c == 0) break; }
        }
        if (best < 0) { ok = false; break; }
        dstpos[s0 + k] = (best >> 1) * 32 + (best & 1) * 16 + fill[best];
        rmask[best] |= (unsigned short)(1u << rc);
        cmask[best] |= (unsigned short)(1u << cc);
        fill[best]++;
        used |= 1u << (best >> 1);
      }
    }
    if (!ok) {
      for (int k = 0; k < cnt; ++k) {
        const int grp = k % G, chunk = k / G;
        dstpos[s0 + k] = grp * 32 + (((chunk & 1) << 4) | (chunk >> 1));
      }
    }
  }
}

__global__ void __launch_bounds__(kThreads)
k_tl_scatter(const unsigned *__restrict__ skey, const unsigned *__restrict__ sval, long long total, long long nnz1,
             const int *__restrict__ sstart, const int *__restrict__ gbase, const int *__restrict__ rowof, CsrDev M1,
             CsrDev M2, unsigned *__restrict__ pk, double *__restrict__ val, const int *__restrict__ dstpos) {
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < total; p += (long long)gridDim.x * blockDim.x) {
    const unsigned seg = skey[p] >> 4;
    const unsigned src = sval[p];
    const int k = (int)(p - sstart[seg]);
    const int g0 = gbase[seg];
    int grp, lane;
    if (dstpos) {  // dealt by k_tl_deal
      const int d = dstpos[p];
      grp = d >> 5; lane = d & 31;
    } else {       // round-1 dealing: entry k -> group k mod G, lane k div G
      const int G = gbase[seg + 1] - g0;
      grp = k % G; lane = k / G;
    }
    int col;
    double v;
    if ((long long)src < nnz1) { col = M1.idx[src]; v = M1.val[src]; }
    else { col = M2.idx[src - nnz1]; v = M2.val[src - nnz1]; }
    const int rl = rowof[src] % kTR;
    const size_t dst = ((size_t)g0 + grp) * 32 + lane;
    pk[dst] = ((unsigned)rl << kTRowShift) | (grp == 0 ? kTFlag : 0u) | (unsigned)(col % kTC);
    val[dst] = v;
  }
}

// padding-only segments (the warp has no entry in a column bin other warps use) still mark the bin
__global__ void __launch_bounds__(kThreads)
k_tl_flag_empty(const int *__restrict__ cnt, const int *__restrict__ gbase, long long nseg, unsigned *__restrict__ pk) {
  for (long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x; s < nseg; s += (long long)gridDim.x * blockDim.x)
    if (cnt[s] == 0 && gbase[s + 1] - gbase[s] == 1) pk[(size_t)gbase[s] * 32] |= kTFlag;
}

__global__ void __launch_bounds__(kThreads) k_tl_iota(unsigned *__restrict__ x, long long n) {
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x)
    x[t] = (unsigned)t;
}

struct HostItem {
  int rb, first, last;  // range in the row bin's active list
  int piece;
  double cost;
};

// SCS_B200_TILED_DIRECT=0 keeps short-row bins in the tiled format (tests / comparison runs)
bool force_tiled_only() {
  const char *e = getenv("SCS_B200_TILED_DIRECT");
  return e && *e && atoi(e) == 0;
}

}  // namespace

// Host plan of a tiled operator (no device work): which row bins are tiled, how their (row bin, active
// column bin) cells are cut into one contiguous, cost-balanced range per CTA, and where every piece parks
// its partial sums.  hg[rb * ncb + cb] = groups of the cell (summed over the warps), p1 / p2 = row pointers
// of the source matrices (p2 may be null).
void tiled_plan_host(int nrows, int ncb, const int *hg, const int *p1, const int *p2, int sms, bool tiled_only,
                     TiledPlan &P) {
  const int nrb = (nrows + kTR - 1) / kTR;
  struct Atom { int rb, cell; double cost; };  // one active column bin of a tiled row bin
  std::vector<std::vector<int>> act((size_t)nrb);
  std::vector<Atom> atoms;
  std::vector<char> is_direct((size_t)nrb, 0);
  const double kCellFixed = 8.0 * kTC, kSlot = 12.0, kItemFixed = 16.0 * kTR, kCellFloor = 70000.0;
  for (int rb = 0; rb < nrb; ++rb) {
    const int r0 = rb * kTR, r1 = std::min(nrows, r0 + kTR);
    long long bn = 0;
    int maxlen = 0;
    for (int r = r0; r < r1; ++r) {
      const int len = (p1[r + 1] - p1[r]) + (p2 ? p2[r + 1] - p2[r] : 0);
      bn += len;
      maxlen = std::max(maxlen, len);
    }
    if (bn == 0 || (!tiled_only && bn <= 6ll * (r1 - r0) && maxlen <= 32)) {
      is_direct[rb] = 1;
      continue;
    }
    for (int cb = 0; cb < ncb; ++cb)
      if (hg[(size_t)rb * ncb + cb] > 0) act[rb].push_back(cb);
    const int na = (int)act[rb].size();
    for (int a = 0; a < na; ++a) {
      // a column bin costs its stream + its x-slice, and never less than the latency of one slice hand-over
      double cst = std::max(kSlot * 32.0 * hg[(size_t)rb * ncb + act[rb][a]] + kCellFixed, kCellFloor);
      atoms.push_back(Atom{rb, a, cst});
    }
  }
  double ctot = 0.0;
  for (const Atom &a : atoms) ctot += a.cost;
  const int natoms = (int)atoms.size();
  const int ncta = std::max(1, std::min(sms, natoms));
  // cut points: atoms [cut[b], cut[b+1]) go to CTA b
  std::vector<int> cut((size_t)ncta + 1, natoms);
  cut[0] = 0;
  {
    double run = 0.0;
    int b = 1;
    for (int i = 0; i < natoms && b < ncta; ++i) {
      run += atoms[i].cost;
      while (b < ncta && run >= ctot * b / ncta) cut[b++] = i + 1;
    }
    // snap a cut that would leave a sliver of a row bin (fewer than 3 cells) to the bin boundary
    for (int b2 = 1; b2 < ncta; ++b2) {
      int k = cut[b2];
      if (k <= 0 || k >= natoms || atoms[k - 1].rb != atoms[k].rb) continue;
      int first = k, last = k;
      while (first > 0 && atoms[first - 1].rb == atoms[k].rb) --first;
      while (last < natoms && atoms[last].rb == atoms[k].rb) ++last;
      if (k - first < 3) k = first;
      else if (last - k < 3) k = last;
      cut[b2] = std::max(k, cut[b2 - 1]);
    }
    for (int b2 = 1; b2 <= ncta; ++b2) cut[b2] = std::max(cut[b2], cut[b2 - 1]);
  }
  // items: maximal runs of one row bin inside a CTA's range
  std::vector<HostItem> items;
  std::vector<std::vector<int>> mine((size_t)ncta);
  std::vector<int> pieces_of((size_t)nrb, 0);
  for (int b = 0; b < ncta; ++b) {
    int i = cut[b];
    while (i < cut[b + 1]) {
      int j = i + 1;
      while (j < cut[b + 1] && atoms[j].rb == atoms[i].rb) ++j;
      HostItem hi;
      hi.rb = atoms[i].rb; hi.first = atoms[i].cell; hi.last = atoms[j - 1].cell + 1;
      hi.piece = pieces_of[hi.rb]++;
      hi.cost = kItemFixed;
      for (int k = i; k < j; ++k) hi.cost += atoms[k].cost;
      mine[b].push_back((int)items.size());
      items.push_back(hi);
      i = j;
    }
  }
  int pslots = 0;
  int pieces_max = 1;
  std::vector<int2> h_binfo((size_t)nrb);
  for (int rb = 0; rb < nrb; ++rb) {
    h_binfo[rb] = make_int2(pslots, pieces_of[rb]);  // pieces == 0: short-row bin (or no rows at all)
    pslots += pieces_of[rb];
    pieces_max = std::max(pieces_max, pieces_of[rb]);
  }
  const int nitems = (int)items.size();
  P.cta_cost.assign((size_t)ncta, 0.0);
  P.cta_items.assign((size_t)ncta, 0);
  for (int b = 0; b < ncta; ++b) {
    for (int i : mine[b]) P.cta_cost[b] += items[i].cost;
    P.cta_items[b] = (int)mine[b].size();
  }
  std::vector<TItem> &h_items = P.items;
  std::vector<int> &h_cta_off = P.cta_off, &h_seq = P.seq, &h_seq_off = P.seq_off;
  h_items.clear(); h_seq.clear();
  h_cta_off.assign(1, 0); h_seq_off.assign(1, 0);
  for (int b = 0; b < ncta; ++b) {
    for (int i : mine[b]) {
      const HostItem &hi = items[i];
      TItem t;
      t.rb = hi.rb;
      t.slot = h_binfo[hi.rb].x + hi.piece;
      t.a0 = (int)h_seq.size();
      for (int a = hi.first; a < hi.last; ++a) h_seq.push_back(act[hi.rb][a]);
      t.a1 = (int)h_seq.size();
      h_items.push_back(t);
    }
    h_cta_off.push_back((int)h_items.size());
    h_seq_off.push_back((int)h_seq.size());
  }
  const int nitems_total = (int)P.items.size();
  (void)nitems_total;
  P.ncta = ncta; P.pslots = pslots; P.pieces_max = pieces_max;
  P.binfo = h_binfo;
  P.ndirect = (int)std::count(is_direct.begin(), is_direct.end(), (char)1);
}

// test hook (host only): the plan for a synthetic cell map.  items_out: 5 ints per item {cta, row bin, slot,
// first index into seq_out, one past the last}; binfo_out: 2 ints per row bin {first slot, pieces};
// cost_out: one double per CTA.  Returns the number of items, or -1 when a buffer is too small; *ncta_out
// and *nseq_out receive the CTA and sequence counts.
extern "C" scs_int scs_b200_tiled_plan(scs_int nrows, scs_int ncb, const scs_int *hg, const scs_int *p1, const scs_int *p2,
                                       scs_int sms, scs_int *items_out, scs_int items_cap, scs_int *seq_out, scs_int seq_cap,
                                       scs_int *binfo_out, double *cost_out, scs_int *ncta_out, scs_int *nseq_out) {
  if (nrows <= 0 || ncb <= 0 || !hg || !p1 || sms <= 0 || !items_out || !seq_out || !binfo_out || !cost_out) return -1;
  TiledPlan P;
  tiled_plan_host(nrows, ncb, hg, p1, p2, sms, false, P);
  const int nitems = (int)P.items.size(), nrb = (nrows + kTR - 1) / kTR;
  if (nitems > items_cap || (int)P.seq.size() > seq_cap) return -1;
  for (int b = 0; b < P.ncta; ++b)
    for (int i = P.cta_off[b]; i < P.cta_off[b + 1]; ++i) {
      items_out[5 * i + 0] = b; items_out[5 * i + 1] = P.items[i].rb; items_out[5 * i + 2] = P.items[i].slot;
      items_out[5 * i + 3] = P.items[i].a0; items_out[5 * i + 4] = P.items[i].a1;
    }
  for (size_t k = 0; k < P.seq.size(); ++k) seq_out[k] = P.seq[k];
  for (int rb = 0; rb < nrb; ++rb) { binfo_out[2 * rb] = P.binfo[rb].x; binfo_out[2 * rb + 1] = P.binfo[rb].y; }
  for (int b = 0; b < P.ncta; ++b) cost_out[b] = P.cta_cost[b];
  if (ncta_out) *ncta_out = P.ncta;
  if (nseq_out) *nseq_out = (int)P.seq.size();
  return nitems;
}

// tile geometry of this build (csrc/tiled.cuh, -DB200_TILED_CFG): rows per row bin, columns per column bin,
// warps per CTA, x-slice stages
extern "C" void scs_b200_tiled_geometry(scs_int out[4]) {
  out[0] = kTR; out[1] = kTC; out[2] = kTW; out[3] = kTStages;
}

void TiledOp::destroy() {
  dev_free(d.items); dev_free(d.cta_off); dev_free(d.seq); dev_free(d.cta_seq_off); dev_free(d.gbase);
  dev_free(d.pk); dev_free(d.val); dev_free(d.partial); dev_free(d.binfo); dev_free(d.prof);
  dev_free(d.dbins); dev_free(d.ydir);
  d.ndbins = 0;
  has_tiled = false;
  ok = false;
}

int TiledOp::build(Ctx &c, const CsrDev &m1, const CsrDev *m2, bool force) {
  ok = false;
  const int nrows = m1.nrows;
  const long long nnz1 = m1.nnz, nnz2 = m2 ? m2->nnz : 0, total = nnz1 + nnz2;
  nnz = total;
  if (nrows <= 0 || total <= 0 || total > 0x7fffffffLL) return 0;
  // Small matrices are better off on the row engine: too few (row bin, column bin) cells to keep 148 CTAs
  // busy, and two launches per product.  Measured on the LASSO family (profiles/r1w_tiled_threshold.txt):
  // 6.6 M entries: row engine 25 % faster end to end; 26.5 M: tiled 1.3x (A) / 0.96x (A' | P); 53 M: 1.58x /
  // 1.31x; 106 M: 1.36x / 1.36x.  SCS_B200_TILED_MIN_NNZ overrides the threshold (tests).
  long long min_nnz = 16ll << 20;
  if (const char *e = getenv("SCS_B200_TILED_MIN_NNZ")) min_nnz = atoll(e);
  if (!force && total < min_nnz) return 0;
  const int nrb = (nrows + kTR - 1) / kTR;
  const int ncbA = (m1.ncols + kTC - 1) / kTC > 0 ? (m1.ncols + kTC - 1) / kTC : 1;
  const int ncbB = m2 ? (m2->ncols + kTC - 1) / kTC : 0;
  const int ncb = ncbA + ncbB;
  const long long nseg = (long long)nrb * kTW * ncb;
  if (nseg > (1ll << 26)) return 0;
  cudaStream_t st = c.stream;
  const int grid = c.grid_ew();
  int *cnt = nullptr, *maxrow = nullptr, *ng = nullptr, *gcnt = nullptr, *sstart = nullptr, *rowof = nullptr;
  unsigned *key = nullptr, *key2 = nullptr, *sv = nullptr, *sv2 = nullptr;
  void *tmp = nullptr;
  int rc = -1;
  const bool verbose = getenv("SCS_B200_TILED_VERBOSE") != nullptr;
  auto tp0 = std::chrono::steady_clock::now();
  auto lap = [&](const char *what) {  // stage clock of the builder (diagnostics only)
    if (!verbose) return;
    cudaStreamSynchronize(st);
    auto t = std::chrono::steady_clock::now();
    fprintf(stderr, "tiled build: %-12s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(t - tp0).count());
    tp0 = t;
  };
  do {
    if (dev_alloc_zero(&cnt, (size_t)nseg + 1, st) || dev_alloc_zero(&maxrow, (size_t)nseg + 1, st) ||
        dev_alloc_zero(&ng, (size_t)nseg + 1, st) || dev_alloc(&gcnt, (size_t)nrb * ncb) ||
        dev_alloc(&sstart, (size_t)nseg + 1) || dev_alloc(&d.gbase, (size_t)nseg + 1) ||
        dev_alloc(&rowof, (size_t)total) || dev_alloc(&key, (size_t)total) || dev_alloc(&key2, (size_t)total) ||
        dev_alloc(&sv, (size_t)total) || dev_alloc(&sv2, (size_t)total))
      break;
    lap("alloc");
    k_tl_count<<<grid, kThreads, 0, st>>>(m1, 0, 0, ncb, key, rowof, cnt, maxrow);
    if (m2 && nnz2 > 0) k_tl_count<<<grid, kThreads, 0, st>>>(*m2, ncbA, nnz1, ncb, key, rowof, cnt, maxrow);
    k_tl_groups<<<grid, kThreads, 0, st>>>(cnt, maxrow, nrb, ncb, ng, gcnt);
    c.launches += 3;
    lap("count");
    size_t tb1 = 0, tb2 = 0, tb3 = 0;
    if (cub::DeviceScan::ExclusiveSum(nullptr, tb1, ng, d.gbase, (int)(nseg + 1), st) != cudaSuccess) break;
    int bits = 1;
    while ((1ll << bits) < nseg * 16 && bits < 32) ++bits;  // key = segment * 16 + bank class
    if (cub::DeviceRadixSort::SortPairs(nullptr, tb2, key, key2, sv, sv2, (int)total, 0, bits, st) != cudaSuccess) break;
    tb3 = tb1 > tb2 ? tb1 : tb2;
    if (cudaMalloc(&tmp, tb3 ? tb3 : 16) != cudaSuccess) break;
    if (cub::DeviceScan::ExclusiveSum(tmp, tb3, ng, d.gbase, (int)(nseg + 1), st) != cudaSuccess) break;
    if (cub::DeviceScan::ExclusiveSum(tmp, tb3, cnt, sstart, (int)(nseg + 1), st) != cudaSuccess) break;
    int groups_total = 0;
    if (cudaMemcpyAsync(&groups_total, d.gbase + nseg, sizeof(int), cudaMemcpyDeviceToHost, st) != cudaSuccess) break;
    std::vector<int> hg((size_t)nrb * ncb);
    if (cudaMemcpyAsync(hg.data(), gcnt, sizeof(int) * hg.size(), cudaMemcpyDeviceToHost, st) != cudaSuccess) break;
    if (cudaStreamSynchronize(st) != cudaSuccess) break;
    lap("scan");
    slots = (long long)groups_total * 32;
    if (groups_total <= 0 || (!force && (double)slots > 1.35 * (double)total)) { rc = 0; break; }  // too much padding
    // ---- entry stream
    if (dev_alloc(&d.pk, (size_t)slots) || dev_alloc(&d.val, (size_t)slots)) break;
    k_tl_iota<<<grid, kThreads, 0, st>>>(sv, total);
    if (cub::DeviceRadixSort::SortPairs(tmp, tb3, key, key2, sv, sv2, (int)total, 0, bits, st) != cudaSuccess) break;
    k_tl_pad<<<grid, kThreads, 0, st>>>(d.pk, d.val, slots);
    // SCS_B200_TILED_DEAL: "greedy" (default) / "perm" (plain dealing, half-warp-aware lanes) / "r1" (round-1 dealing)
    const char *deal_env = getenv("SCS_B200_TILED_DEAL");
    const int deal = (deal_env && !strcmp(deal_env, "r1")) ? 0 : ((deal_env && !strcmp(deal_env, "perm")) ? 1 : 2);
    int *dstpos = nullptr;
    if (deal > 0) {
      if (dev_alloc(&dstpos, (size_t)total)) break;
      const long long blocks = (nseg + 127) / 128;
      k_tl_deal<<<(unsigned)std::min<long long>(blocks, 1 << 20), 128, 0, st>>>(sv2, nnz1, sstart, d.gbase, nseg, rowof, m1,
                                                                                m2 ? *m2 : m1, dstpos, deal == 2 ? 1 : 0);
      c.launches++;
    }
    k_tl_scatter<<<grid, kThreads, 0, st>>>(key2, sv2, total, nnz1, sstart, d.gbase, rowof, m1, m2 ? *m2 : m1, d.pk, d.val,
                                            dstpos);
    if (dstpos) { cudaStreamSynchronize(st); dev_free(dstpos); }
    k_tl_flag_empty<<<grid, kThreads, 0, st>>>(cnt, d.gbase, nseg, d.pk);
    c.launches += 4;
    lap("sort+scatter");
    // ---- host plan.  Row bins whose rows are all short (identity / bound blocks: a couple of entries per
    // row) are not worth an x-slice per column bin: they become "direct" items, handled row by row
    // from the CSR arrays inside the same kernel.  The rest is a sequence of (row bin, active column
    // bin) cells; the whole sequence is cut into one contiguous range of equal modelled cost per CTA, so
    // a tall row bin is shared by the two or three CTAs whose ranges meet in it.
    std::vector<int> p1((size_t)nrows + 1), p2;
    if (cudaMemcpyAsync(p1.data(), m1.ptr, sizeof(int) * p1.size(), cudaMemcpyDeviceToHost, st) != cudaSuccess) break;
    if (m2) {
      p2.resize((size_t)nrows + 1);
      if (cudaMemcpyAsync(p2.data(), m2->ptr, sizeof(int) * p2.size(), cudaMemcpyDeviceToHost, st) != cudaSuccess) break;
    }
    if (cudaStreamSynchronize(st) != cudaSuccess) break;
    TiledPlan plan;
    tiled_plan_host(nrows, ncb, hg.data(), p1.data(), m2 ? p2.data() : nullptr, c.sms, force_tiled_only(), plan);
    const std::vector<TItem> &h_items = plan.items;
    const std::vector<int> &h_cta_off = plan.cta_off, &h_seq = plan.seq, &h_seq_off = plan.seq_off;
    const std::vector<int2> &h_binfo = plan.binfo;
    const int ncta = plan.ncta, pslots = plan.pslots, nitems = (int)h_items.size();
    pieces_max = plan.pieces_max;
    has_tiled = nitems > 0;
    cta_cost = plan.cta_cost;
    cta_items = plan.cta_items;
    if (dev_alloc(&d.items, h_items.size()) || dev_alloc(&d.cta_off, h_cta_off.size()) ||
        dev_alloc(&d.seq, h_seq.size()) || dev_alloc(&d.cta_seq_off, h_seq_off.size()) ||
        dev_alloc(&d.partial, (size_t)std::max(pslots, 1) * kTR) || dev_alloc(&d.binfo, (size_t)nrb) ||
        dev_alloc_zero(&d.prof, (size_t)3 * ncta, st))
      break;
    if (cudaMemcpyAsync(d.items, h_items.data(), sizeof(TItem) * h_items.size(), cudaMemcpyHostToDevice, st) != cudaSuccess ||
        cudaMemcpyAsync(d.cta_off, h_cta_off.data(), sizeof(int) * h_cta_off.size(), cudaMemcpyHostToDevice, st) != cudaSuccess ||
        cudaMemcpyAsync(d.seq, h_seq.data(), sizeof(int) * h_seq.size(), cudaMemcpyHostToDevice, st) != cudaSuccess ||
        cudaMemcpyAsync(d.cta_seq_off, h_seq_off.data(), sizeof(int) * h_seq_off.size(), cudaMemcpyHostToDevice, st) != cudaSuccess ||
        cudaMemcpyAsync(d.binfo, h_binfo.data(), sizeof(int2) * h_binfo.size(), cudaMemcpyHostToDevice, st) != cudaSuccess)
      break;
    {  // short-row bins: their list, and the raw-product scratch of the side-stream pass
      std::vector<int> db;
      for (int rb = 0; rb < nrb; ++rb) {
        const int r0 = rb * kTR, r1 = std::min(nrows, r0 + kTR);
        const bool empty = p1[(size_t)r1] == p1[(size_t)r0] && (!m2 || p2[(size_t)r1] == p2[(size_t)r0]);
        if (h_binfo[(size_t)rb].y == 0 && !empty) db.push_back(rb);
      }
      // "1": multiply the short rows out on a side stream in the shadow of the streaming kernel.  Off by default:
      // measured on the bench workload it does not pay (0.436 -> 0.440 ms per product, profiles/r2c_spmv_variants.txt);
      // the streaming kernel is latency-bound and the extra traffic slows it by as much as the epilogue pass gains
      const char *e = getenv("SCS_B200_TILED_SIDE");
      const bool side_on = e && atoi(e) == 1;
      d.ndbins = 0;
      if (side_on && nitems > 0 && !db.empty()) {
        if (dev_alloc(&d.dbins, db.size()) || dev_alloc_zero(&d.ydir, (size_t)nrows, st) ||
            cudaMemcpyAsync(d.dbins, db.data(), sizeof(int) * db.size(), cudaMemcpyHostToDevice, st) != cudaSuccess)
          break;
        d.ndbins = (int)db.size();
      }
    }
    {
      const char *e = getenv("SCS_B200_TILED_U");  // groups per register buffer of the streaming kernel: 8 or 6
      variant = (e && atoi(e) == 6) ? 1 : 0;      // (12 was measured 15 % slower: 122 registers, profiles/r2c_*)
      const char *e2 = getenv("SCS_B200_TILED_EB");
      // epilogue pass (SCS_B200_TILED_EB): 1 = chunked / entry-parallel short rows, 0 = the same with 8 entries per thread
      // batched, 4 / 8 = row-parallel with that many rows in flight per thread
      epi_eb = (e2 && *e2) ? atoi(e2) : -1;
      // default: two-source operators (Gp) take the chunked kernel, single-source ones (A p) the row-parallel one;
      // the forms differ by ~1 % of a product (profiles/r2q_spmv_epilogue_forms.txt)
      if (epi_eb != 0 && epi_eb != 1 && epi_eb != 4 && epi_eb != 8) epi_eb = m2 ? 1 : 4;
    }
    if (cudaStreamSynchronize(st) != cudaSuccess || cudaGetLastError() != cudaSuccess) break;
    lap("plan+upload");
    if (verbose)
      fprintf(stderr, "tiled build: rows %d nnz %lld slots %lld (x%.3f) row bins %d col bins %d items %d ctas %d max pieces %d direct bins %d\n",
              nrows, total, slots, (double)slots / (double)total, nrb, ncb, nitems, ncta, pieces_max, plan.ndirect);
    d.nrows = nrows; d.nrb = nrb; d.ncb = ncb; d.ncbA = ncbA;
    d.len1 = m1.ncols; d.len2 = m2 ? m2->ncols : 0;
    d.nitems = nitems; d.ncta = ncta;
    this->m1 = m1; has2 = m2 != nullptr; this->m2 = m2 ? *m2 : m1;
    ok = true;
    rc = 0;
  } while (0);
  cudaStreamSynchronize(st);
  if (tmp) cudaFree(tmp);
  dev_free(cnt); dev_free(maxrow); dev_free(ng); dev_free(gcnt); dev_free(sstart); dev_free(rowof);
  dev_free(key); dev_free(key2); dev_free(sv); dev_free(sv2);
  if (!ok) destroy();
  if (rc != 0) fprintf(stderr, "libscsb200: building the tiled SpMV format failed (%s)\n", cudaGetErrorString(cudaGetLastError()));
  return rc;
}

}  // namespace b200
