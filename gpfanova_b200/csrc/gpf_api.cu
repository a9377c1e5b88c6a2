// C ABI of gpfanova_b200 (include/gpfanova_b200.h): host-side context, launch logic and the sweep driver
// (SamplerContainer.sample/_sample/store, sample/container.py:29-92; sampler order of base.py:57-84).
// There is no CPU path in this file: every compute entry point launches sm_100a kernels or fails.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/gpfanova_b200.h"
#include "gpf_kernels.cuh"
#include "gpf_cond.cuh"
#include "gpf_ops.cuh"
#include "gpf_schur_args.cuh"

using namespace gpf;

namespace {

thread_local std::string g_err;   // for the stateless operators (no ctx)

struct Samp { int kind; int arg; };   // kind 0: slice on hyper `arg`; kind 1: function `arg`

size_t smem_limit() { return 227 * 1024; }

template <class K>
cudaError_t allow_smem(K kernel, size_t bytes) {
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

}  // namespace

struct gpf_ctx {
  CtxDev dev{};
  int device = 0;
  int sms = 148;
  int grid = 1;
  cudaStream_t stream = nullptr;
  std::string err;
  std::vector<double> slice_w;
  std::vector<int> slice_m;
  std::vector<int> group_of_fn;
  std::vector<std::vector<int>> groups;
  std::vector<Samp> order;
  std::vector<char> lb_valid;         // per group: c.LB holds jitchol(K_g + offset I) at the current hyper-parameters
  bool toeplitz = false;
  bool centro_ok = false;
  const double* d_bhat = nullptr;     // least-squares functions for the observation likelihood (balanced designs)
  bool indep = false;                 // X^T X diagonal and nothing missing: function draws of a sweep are independent
  int* d_fn_all = nullptr;            // all functions in sampler order / their sampler ids (device)
  unsigned* d_sid_all = nullptr;
  int ncls = 0;                       // classes of functions sharing their conditional precision (indep designs)
  int *d_cls_off = nullptr, *d_cls_fn = nullptr, *d_gcls_off = nullptr, *d_gorder = nullptr;
  int *d_cls1_off = nullptr, *d_gcls1_off = nullptr;   // option 2 = 0: every function its own class
  unsigned* d_cls_sid = nullptr;
  double* d_kinv_tmp = nullptr;       // scratch of the K_inv operator (gpf_group_kinv with an output buffer)
  bool schurfn_ok = false;            // uniform grid, balanced design, n <= 256, groups of at most 4: O(n^2) function step
  int use_schurfn = 1;
  int pipelines = 1;                  // balanced sweeps on the Schur kernels: functions -> slices of each group on its own stream;
                                      // 2: the slice kernels also store the rows and run the next y_sigma step (fold) -- fewer
                                      // launches, but measured slower on the device (5.82 vs 5.71 ms): one warp per chain does
                                      // at the end of the longest units what k_obs does for all chains at once
  int cohorts = 1;                    // ... and the chains in this many independent cohorts, each with its own streams (1 or 2;
                                      // two were SLOWER at every chain count: more kernel variants per SM, smaller grids)
  struct Lane {                       // streams / events of the second cohort (the first uses stream, aux, ev_fork, ev_join)
    cudaStream_t st = nullptr, aux[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t fork = nullptr, join[4] = {nullptr, nullptr, nullptr, nullptr}, begin = nullptr, end = nullptr;
  } lane2;
  std::vector<int> h_gorder;
  int use_sweepk = 0;                 // whole sweeps in one launch, no device-wide barriers (k_chain_sweeps) where it applies;
                                      // off by default: its instruction footprint (function step + three slice variants in
                                      // flight on every SM) misses the 32 KB instruction cache -- measured 4x slower
  int* d_swk = nullptr;               // its per-chain progress words: prog[C], done[C], err
  unsigned* d_sid_sigma = nullptr;    // sampler id of prior{g}_sigma per group
  std::vector<unsigned> h_sid_sigma;
  double* d_slice_w = nullptr;
  int* d_slice_m = nullptr;
  std::vector<int> schur_warps;       // per group: warps per CTA of the Schur slice kernel (0: dense engine)
  int use_schur = 1;
  int batch = 1;                       // batch independent function draws / K_inv of all groups into one launch
  int share = 1;                       // one factorisation per class of functions with identical conditional precision
  cudaStream_t aux[4] = {nullptr, nullptr, nullptr, nullptr};   // the groups' slice kernels of a batched sweep run concurrently
  cudaEvent_t ev_fork = nullptr, ev_join[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev_user[4] = {nullptr, nullptr, nullptr, nullptr};   // gpf_record_event / gpf_wait_event
  cudaStream_t copy = nullptr;         // device -> host history copies run beside the kernels of the next sweeps
  cudaEvent_t ev_copy = nullptr;
  std::vector<void*> allocs;
  size_t smem_kinv = 0, smem_kinv2 = 0, smem_fn[3] = {0, 0, 0}, smem_fact[3] = {0, 0, 0}, smem_fused[3] = {0, 0, 0};
  std::vector<size_t> smem_prior;
  unsigned long long launches = 0;
  int profile = 0;                     // 0 off, 1 per-kernel + total CUDA-event timing, 2 total only
  double ms[6] = {0, 0, 0, 0, 0, 0};
  struct Ev { cudaEvent_t a, b; int cat; };
  std::vector<Ev> evs;
  std::vector<cudaEvent_t> ev_pool;
};

#define CK(ctx, call)                                                                        \
  do {                                                                                       \
    cudaError_t e_ = (call);                                                                 \
    if (e_ != cudaSuccess) {                                                                 \
      char buf_[512];                                                                        \
      snprintf(buf_, sizeof buf_, "%s:%d: %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
      (ctx)->err = buf_;                                                                     \
      return GPF_ERR_CUDA;                                                                   \
    }                                                                                        \
  } while (0)

#define CKS(call)                                                                            \
  do {                                                                                       \
    cudaError_t e_ = (call);                                                                 \
    if (e_ != cudaSuccess) {                                                                 \
      char buf_[512];                                                                        \
      snprintf(buf_, sizeof buf_, "%s:%d: %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
      g_err = buf_;                                                                          \
      return GPF_ERR_CUDA;                                                                   \
    }                                                                                        \
  } while (0)

namespace {

struct Timer {   // optional per-launch CUDA-event bracket (gpf_set_profile)
  gpf_ctx* c;
  int cat;
  cudaEvent_t a = nullptr, b = nullptr;
  Timer(gpf_ctx* ctx, int category) : c(ctx), cat(category) {
    if (c->profile != 1) return;
    auto get = [&]() {
      cudaEvent_t e;
      if (!c->ev_pool.empty()) { e = c->ev_pool.back(); c->ev_pool.pop_back(); }
      else cudaEventCreate(&e);
      return e;
    };
    a = get(); b = get();
    cudaEventRecord(a, c->stream);
  }
  ~Timer() {
    if (c->profile != 1) return;
    cudaEventRecord(b, c->stream);
    c->evs.push_back({a, b, cat});
  }
};

void collect_profile(gpf_ctx* c) {
  for (auto& e : c->evs) {
    float t = 0.f;
    cudaEventSynchronize(e.b);
    cudaEventElapsedTime(&t, e.a, e.b);
    c->ms[e.cat] += t;
    c->ev_pool.push_back(e.a);
    c->ev_pool.push_back(e.b);
  }
  c->evs.clear();
}

template <class T>
int dev_alloc(gpf_ctx* c, T** p, size_t count) {
  void* q = nullptr;
  CK(c, cudaMalloc(&q, std::max<size_t>(count, 1) * sizeof(T)));
  c->allocs.push_back(q);
  *p = (T*)q;
  return GPF_OK;
}

int mode_of(const gpf_ctx* c) { return c->dev.centro ? 2 : (c->toeplitz ? 1 : 0); }

// Kernel.K_inv of group g into the scratch buffer (inspection operator; the sweep itself never forms K_inv)
int launch_kinv(gpf_ctx* c, int g) {
  Timer t(c, 1);
  const size_t tstride = (size_t)plt_tiles(c->dev.NT) * TILE;
  if (!c->d_kinv_tmp) {
    void* q = nullptr;
    CK(c, cudaMalloc(&q, (size_t)c->dev.C * tstride * sizeof(double)));
    c->allocs.push_back(q);
    c->d_kinv_tmp = (double*)q;
  }
  CtxDev dev = c->dev;
  dev.Kinv = c->d_kinv_tmp - (size_t)g * c->dev.C * tstride;     // the kernel indexes (group, chain)
  CK(c, cudaMemsetAsync(c->dev.sched, 0, sizeof(int), c->stream));
  const int grid = std::min(c->dev.C, 2 * c->sms);
  if (c->dev.centro) k_group_kinv<2><<<grid, NTHR, c->smem_kinv2, c->stream>>>(dev, g, 1);
  else if (c->toeplitz) k_group_kinv<1><<<grid, NTHR, c->smem_kinv, c->stream>>>(dev, g, 1);
  else k_group_kinv<0><<<grid, NTHR, c->smem_kinv, c->stream>>>(dev, g, 1);
  ++c->launches;
  CK(c, cudaGetLastError());
  return GPF_OK;
}

// jitchol(K_g + offset I) of groups g .. g+count-1 into c.LB (sequential sweeps: the function steps read it)
int launch_factor(gpf_ctx* c, int g, int count = 1) {
  Timer t(c, 1);
  CK(c, cudaMemsetAsync(c->dev.sched, 0, sizeof(int), c->stream));
  const int grid = std::min(c->dev.C * count, 2 * c->sms);
  const int mode = mode_of(c);
  if (mode == 2) k_group_factor<2><<<grid, NTHR, c->smem_fact[2], c->stream>>>(c->dev, g, count);
  else if (mode == 1) k_group_factor<1><<<grid, NTHR, c->smem_fact[1], c->stream>>>(c->dev, g, count);
  else k_group_factor<0><<<grid, NTHR, c->smem_fact[0], c->stream>>>(c->dev, g, count);
  ++c->launches;
  CK(c, cudaGetLastError());
  for (int q = g; q < g + count; ++q) c->lb_valid[q] = 1;
  return GPF_OK;
}

int launch_function_kernel(gpf_ctx* c, int grid, int fn, unsigned sid, const int* fl, const unsigned* sl, int nfn, unsigned sweep,
                           const double* z, FnOut out) {
  const int mode = mode_of(c);
  if (mode == 2) k_function<2><<<grid, NTHR, c->smem_fn[2], c->stream>>>(c->dev, fn, sid, fl, sl, nfn, sweep, z, out);
  else if (mode == 1) k_function<1><<<grid, NTHR, c->smem_fn[1], c->stream>>>(c->dev, fn, sid, fl, sl, nfn, sweep, z, out);
  else k_function<0><<<grid, NTHR, c->smem_fn[0], c->stream>>>(c->dev, fn, sid, fl, sl, nfn, sweep, z, out);
  ++c->launches;
  CK(c, cudaGetLastError());
  return GPF_OK;
}

int launch_function(gpf_ctx* c, int fn, unsigned sampler_id, unsigned sweep, const double* z, FnOut out) {
  const int g = c->group_of_fn[fn];
  if (!c->lb_valid[g]) {
    int rc = launch_factor(c, g);
    if (rc) return rc;
  }
  Timer t(c, 2);
  CK(c, cudaMemsetAsync(c->dev.sched, 0, sizeof(int), c->stream));
  return launch_function_kernel(c, c->grid, fn, sampler_id, nullptr, nullptr, 1, sweep, z, out);
}

// every function of the model in one launch (only when the draws are independent of each other, c->indep):
// one item per (prior group, chain) -- factor of K_o, then the group's classes (k_group_fused)
bool schurfn_active(const gpf_ctx* c) { return c->schurfn_ok && c->use_schurfn && !c->dev.centro; }

int launch_all_functions(gpf_ctx* c, unsigned sweep) {
  Timer t(c, 2);
  CK(c, cudaMemsetAsync(c->dev.sched, 0, sizeof(int), c->stream));
  if (schurfn_active(c)) {
    // uniform grid: both K_o and C = I + c K_o are Toeplitz -- O(n^2) Schur recursions, one warp per (group, chain)
    SchurFnArgs a{c->d_gorder, c->share ? c->d_gcls_off : c->d_gcls1_off, c->share ? c->d_cls_off : c->d_cls1_off, c->d_cls_fn,
                  c->d_cls_sid, sweep, 0};
    CK(c, schur_launch_functions(c->dev.C * c->dev.P, c->sms, c->stream, c->dev, a));
    ++c->launches;
    return GPF_OK;
  }
  const int grid = std::min(c->dev.C * c->dev.P, 2 * c->sms);
  FusedArgs a{c->d_gorder, c->share ? c->d_gcls_off : c->d_gcls1_off, c->share ? c->d_cls_off : c->d_cls1_off, c->d_cls_fn,
              c->d_cls_sid, sweep};
  const int mode = mode_of(c);
  const size_t sm = c->smem_fused[mode];
  if (mode == 2) k_group_fused<2><<<grid, NTHR, sm, c->stream>>>(c->dev, a);
  else if (mode == 1) k_group_fused<1><<<grid, NTHR, sm, c->stream>>>(c->dev, a);
  else k_group_fused<0><<<grid, NTHR, sm, c->stream>>>(c->dev, a);
  ++c->launches;
  CK(c, cudaGetLastError());
  return GPF_OK;
}

// the O(n^2) Toeplitz path needs a uniform grid and vectors that fit registers or shared memory
bool uses_schur(const gpf_ctx* c, int g) {
  return c->toeplitz && c->use_schur && (c->schur_warps[g] > 0 || (c->dev.n <= 256 && c->groups[g].size() <= 4));
}

// mode: 0/1 evaluate sigma / lengthscale candidates, 2/3 slice step on sigma / lengthscale, 4 fused sigma + lengthscale
int launch_prior(gpf_ctx* c, int g, int mode, const double* cand, double* outv, unsigned sampler_id, unsigned sweep,
                 const double* u, int nu, int* nev, int slot = -1) {
  Timer t(c, 3);
  PriorArgs a{g, mode, cand, outv, c->slice_w[1 + 2 * g], c->slice_w[2 + 2 * g], c->slice_m[1 + 2 * g], c->slice_m[2 + 2 * g],
              sampler_id, sweep, u, nu, nev};
  // slot >= 0: run on auxiliary stream `slot` with its own hand-out counter (concurrent slice kernels of different groups)
  cudaStream_t st = slot >= 0 ? c->aux[slot] : c->stream;
  CtxDev dev = c->dev;
  if (slot >= 0) dev.sched = c->dev.sched + 1 + slot;
  CK(c, cudaMemsetAsync(dev.sched, 0, sizeof(int), st));
  if (uses_schur(c, g)) {
    const int ng = (int)c->groups[g].size();
    if (c->dev.n <= 256 && ng <= 4) {       // vectors in registers
      const int grid = std::min((c->dev.C + 3) / 4, schur_minb(ng <= 2 ? ng : 4) * c->sms);   // 4 warps per CTA
      CK(c, schur_launch_prior(ng <= 2 ? ng : 4, grid, 128, 0, st, dev, a));
    } else {
      const int wpc = c->schur_warps[g];
      const size_t smem = (size_t)wpc * schur_warp_doubles(c->dev.n, ng) * sizeof(double);
      const int grid = std::min((c->dev.C + wpc - 1) / wpc, 2 * c->sms);
      CK(c, schur_launch_prior(0, grid, wpc * 32, smem, st, dev, a));
    }
  } else if (c->toeplitz) k_prior<true><<<c->grid, NTHR, c->smem_prior[g], st>>>(dev, a);
  else k_prior<false><<<c->grid, NTHR, c->smem_prior[g], st>>>(dev, a);
  ++c->launches;
  CK(c, cudaGetLastError());
  if (mode >= 2) c->lb_valid[g] = 0;
  return GPF_OK;
}

// whole sweeps in one launch (gpf_sweepk.cuh): balanced design on a uniform grid, every group on the register-resident
// Schur kernels, reference order of the samplers
bool sweepk_active(const gpf_ctx* c) {
  if (!c->use_sweepk || !c->indep || !c->batch || !schurfn_active(c) || c->dev.bhat == nullptr || c->profile == 1) return false;
  if (c->dev.n > 256) return false;
  for (int g = 0; g < c->dev.P; ++g)
    if (!uses_schur(c, g) || c->groups[g].size() > 4) return false;
  return true;
}

int launch_chain_sweeps(gpf_ctx* c, int n_sweeps, int thin, unsigned sweep0, double* d_history, long long cap, long long* rows) {
  const CtxDev& v = c->dev;
  const int ns = (int)c->order.size();
  if (!c->d_swk) {
    int rc = dev_alloc(c, &c->d_swk, (size_t)2 * v.C + 1);
    if (rc) return rc;
    CK(c, cudaMemset(c->d_swk + 2 * v.C, 0, sizeof(int)));
  }
  if (!c->d_sid_sigma) {
    int rc;
    if ((rc = dev_alloc(c, &c->d_sid_sigma, (size_t)v.P)) || (rc = dev_alloc(c, &c->d_slice_w, (size_t)v.H)) ||
        (rc = dev_alloc(c, &c->d_slice_m, (size_t)v.H)))
      return rc;
  }
  std::vector<unsigned> sid((size_t)v.P, 0u);
  for (int k = 0; k < ns; ++k)
    if (c->order[k].kind == 0 && c->order[k].arg >= 1 && (c->order[k].arg - 1) % 2 == 0) sid[(c->order[k].arg - 1) / 2] = (unsigned)k;
  if (sid != c->h_sid_sigma) {
    CK(c, cudaStreamSynchronize(c->stream));
    CK(c, cudaMemcpy(c->d_sid_sigma, sid.data(), sizeof(unsigned) * sid.size(), cudaMemcpyHostToDevice));
    CK(c, cudaMemcpy(c->d_slice_w, c->slice_w.data(), sizeof(double) * v.H, cudaMemcpyHostToDevice));
    CK(c, cudaMemcpy(c->d_slice_m, c->slice_m.data(), sizeof(int) * v.H, cudaMemcpyHostToDevice));
    c->h_sid_sigma = sid;
  }
  const int step = thin <= 0 ? 1 : thin;
  long long kmax = ((1ll << 31) - 1 - v.C) / ((long long)v.C * v.P) - 1;     // tickets of one launch fit an int
  kmax -= kmax % step;
  if (kmax < step) { c->err = "gpf_sweep: thinning interval too large for one launch"; return GPF_ERR_ARG; }
  const int nld = ((v.n + 7) / 8) * 8;
  const size_t smem = (size_t)4 * SFN_SMEM_ROWS * nld * sizeof(double);
  const int grid = (int)std::min<long long>(((long long)v.C * v.P + 3) / 4, (long long)GPF_SFN_MINB * c->sms);
  int done = 0;
  while (done < n_sweeps) {
    const int k = (int)std::min<long long>(kmax, n_sweeps - done);
    const long long nrows = d_history ? (thin <= 0 ? k : k / step) : 0;
    if (d_history && *rows + nrows > cap) { c->err = "gpf_sweep: history buffer full"; return GPF_ERR_ARG; }
    Timer t(c, 2);
    CK(c, cudaMemsetAsync(v.sched, 0, sizeof(int), c->stream));
    CK(c, cudaMemsetAsync(c->d_swk, 0, sizeof(int) * 2 * v.C, c->stream));
    SweepKArgs a{};
    a.fn = SchurFnArgs{c->d_gorder, c->share ? c->d_gcls_off : c->d_gcls1_off, c->share ? c->d_cls_off : c->d_cls1_off, c->d_cls_fn,
                       c->d_cls_sid, 0u, 0};
    a.sweep0 = sweep0 + (unsigned)done;
    a.n_sweeps = k;
    a.thin = thin;
    a.history = d_history ? d_history + (size_t)(*rows) * v.C * ((size_t)v.H + (size_t)v.f * v.n) : nullptr;
    a.rows_cap = nrows;
    a.sid_sigma = c->d_sid_sigma;
    a.slice_w = c->d_slice_w;
    a.slice_m = c->d_slice_m;
    a.sid_y = 0u;
    a.prog = c->d_swk;
    a.done = c->d_swk + v.C;
    a.err = c->d_swk + 2 * v.C;
    CK(c, schur_launch_chain_sweeps(grid, smem, c->stream, v, a));
    ++c->launches;
    *rows += nrows;
    done += k;
  }
  for (int g = 0; g < v.P; ++g) c->lb_valid[g] = 0;
  return GPF_OK;
}

// Balanced design on a uniform grid, every group on the register-resident Schur kernels: one sweep of the chains
// [c0, c0 + cnt) enqueued on one cohort's streams --  y_sigma | per group on its own stream: functions -> sigma + lengthscale
// slices | join.  Cohorts are independent (chains never interact), so the stage boundaries of one cohort (each waits for
// the slowest of its chains) are filled by the other cohort's kernels.
bool pipe_active(const gpf_ctx* c) {
  if (!c->pipelines || !c->indep || !c->batch || !schurfn_active(c) || c->profile == 1 || c->dev.n > 256) return false;
  for (int g = 0; g < c->dev.P; ++g)
    if (!uses_schur(c, g) || c->groups[g].size() > 4) return false;
  return true;
}

struct FoldPlan { bool on; bool launch_y; int target; double* row; bool next_y; };

int enqueue_schur_sweep(gpf_ctx* c, int q, int c0, int cnt, unsigned sweep, const std::vector<std::pair<int, int>>& todo,
                        const FoldPlan& fp) {
  cudaStream_t st = q ? c->lane2.st : c->stream;
  cudaStream_t* aux = q ? c->lane2.aux : c->aux;
  cudaEvent_t fork = q ? c->lane2.fork : c->ev_fork;
  cudaEvent_t* join = q ? c->lane2.join : c->ev_join;
  CtxDev dv = c->dev;
  dv.C = cnt;
  dv.F += (size_t)c0 * dv.f * dv.n;
  dv.hyper += (size_t)c0 * dv.H;
  dv.status += c0;
  dv.chain_offset += c0;
  int* sched = c->dev.sched + 8 * q;
  if (fp.launch_y) {
    k_obs<<<std::min(cnt, c->sms * 8), NTHR, 0, st>>>(dv, 1, nullptr, nullptr, c->slice_w[0], c->slice_m[0], 0u, sweep, nullptr, 0, nullptr);
    ++c->launches;
    CK(c, cudaGetLastError());
  }
  CK(c, cudaEventRecord(fork, st));
  bool forked[4] = {false, false, false, false};
  for (size_t i = 0; i < todo.size(); ++i) {
    const int k = todo[i].first, g = todo[i].second, slot = g % 4;
    if (!forked[slot]) { CK(c, cudaStreamWaitEvent(aux[slot], fork, 0)); forked[slot] = true; }
    CtxDev d2 = dv;
    d2.sched = sched + 1 + slot;
    int gi = 0;
    while (gi < (int)c->h_gorder.size() && c->h_gorder[gi] != g) ++gi;
    CK(c, cudaMemsetAsync(d2.sched, 0, sizeof(int), aux[slot]));
    SchurFnArgs fa{c->d_gorder + gi, c->share ? c->d_gcls_off : c->d_gcls1_off, c->share ? c->d_cls_off : c->d_cls1_off, c->d_cls_fn,
                   c->d_cls_sid, sweep, 1};
    CK(c, schur_launch_functions(cnt, c->sms, aux[slot], d2, fa));
    CK(c, cudaMemsetAsync(d2.sched, 0, sizeof(int), aux[slot]));
    PriorArgs pa{g, 4, nullptr, nullptr, c->slice_w[1 + 2 * g], c->slice_w[2 + 2 * g], c->slice_m[1 + 2 * g], c->slice_m[2 + 2 * g],
                 (unsigned)k, sweep, nullptr, 0, nullptr};
    if (fp.on) {
      pa.fold_done = c->d_swk + c->dev.C + c0;
      pa.fold_target = fp.target;
      pa.fold_row = fp.row;
      pa.fold_cols = (size_t)dv.H + (size_t)dv.f * dv.n;
      pa.fold_next_y = fp.next_y ? 1 : 0;
      pa.w_y = c->slice_w[0];
      pa.m_y = c->slice_m[0];
    }
    const int ng = (int)c->groups[g].size();
    const int grid = std::min((cnt + 3) / 4, schur_minb(ng <= 2 ? ng : 4) * c->sms);
    CK(c, schur_launch_prior(ng <= 2 ? ng : 4, grid, 128, 0, aux[slot], d2, pa));
    c->launches += 2;
  }
  for (int i = 0; i < 4; ++i)
    if (forked[i]) {
      CK(c, cudaEventRecord(join[i], aux[i]));
      CK(c, cudaStreamWaitEvent(st, join[i], 0));
    }
  return GPF_OK;
}

// n_sweeps sweeps of every chain through enqueue_schur_sweep, rows stored per cohort on the cohort's own stream
int pipelined_sweeps(gpf_ctx* c, int n_sweeps, int thin, unsigned sweep0, double* d_history, long long cap, long long* rows) {
  const CtxDev& v = c->dev;
  const int ns = (int)c->order.size();
  const size_t cols = (size_t)v.H + (size_t)v.f * v.n;
  std::vector<std::pair<int, int>> todo;     // (sampler index of prior{g}_sigma, group), largest groups first
  for (int k = 0; k < ns; ++k)
    if (c->order[k].kind == 0 && c->order[k].arg >= 1 && (c->order[k].arg - 1) % 2 == 0) todo.push_back({k, (c->order[k].arg - 1) / 2});
  std::stable_sort(todo.begin(), todo.end(), [&](const std::pair<int, int>& a, const std::pair<int, int>& b) {
    return c->groups[a.second].size() > c->groups[b.second].size();
  });
  const int ncoh = (c->cohorts >= 2 && v.C >= 64) ? 2 : 1;
  int start[3] = {0, v.C, v.C};
  if (ncoh == 2) start[1] = ((v.C / 2 + 3) / 4) * 4;
  // fold: the slice kernels count the finished groups of every chain; the last one stores the row and runs the next y_sigma
  // step, so only the first sweep of the call launches k_obs and no copy kernels are needed (needs the least-squares form of
  // the observation likelihood, option 4)
  const bool fold = c->pipelines >= 2 && v.bhat != nullptr;
  if (fold) {
    if (!c->d_swk) {
      int rc = dev_alloc(c, &c->d_swk, (size_t)2 * v.C + 1);
      if (rc) return rc;
      CK(c, cudaMemset(c->d_swk + 2 * v.C, 0, sizeof(int)));
    }
    CK(c, cudaMemsetAsync(c->d_swk + v.C, 0, sizeof(int) * v.C, c->stream));     // before lane2.begin is recorded below
  }
  if (ncoh == 2) {      // the second cohort starts behind everything queued so far on the context's stream
    CK(c, cudaEventRecord(c->lane2.begin, c->stream));
    CK(c, cudaStreamWaitEvent(c->lane2.st, c->lane2.begin, 0));
  }
  for (int s = 0; s < n_sweeps; ++s) {
    const bool store = d_history && (thin <= 0 || (s + 1) % thin == 0);
    if (store && *rows >= cap) { c->err = "gpf_sweep: history buffer full"; return GPF_ERR_ARG; }
    for (int q = 0; q < ncoh; ++q) {
      const int c0 = start[q], cnt = start[q + 1] - start[q];
      double* row = store ? d_history + ((size_t)(*rows) * v.C + c0) * cols : nullptr;
      FoldPlan fp{fold, !fold || s == 0, v.P * (s + 1), row, s + 1 < n_sweeps};
      int rc = enqueue_schur_sweep(c, q, c0, cnt, sweep0 + (unsigned)s, todo, fp);
      if (rc) return rc;
      if (store && !fold) {
        cudaStream_t st = q ? c->lane2.st : c->stream;
        CK(c, cudaMemcpy2DAsync(row, cols * sizeof(double), v.hyper + (size_t)c0 * v.H, v.H * sizeof(double), v.H * sizeof(double), cnt,
                                cudaMemcpyDeviceToDevice, st));
        CK(c, cudaMemcpy2DAsync(row + v.H, cols * sizeof(double), v.F + (size_t)c0 * v.f * v.n, (size_t)v.f * v.n * sizeof(double),
                                (size_t)v.f * v.n * sizeof(double), cnt, cudaMemcpyDeviceToDevice, st));
      }
    }
    if (store) ++*rows;
  }
  if (ncoh == 2) {
    CK(c, cudaEventRecord(c->lane2.end, c->lane2.st));
    CK(c, cudaStreamWaitEvent(c->stream, c->lane2.end, 0));
  }
  for (int g = 0; g < v.P; ++g) c->lb_valid[g] = 0;
  return GPF_OK;
}

int launch_slice(gpf_ctx* c, int h, int mode, const double* cand, double* outv, unsigned sampler_id, unsigned sweep,
                 const double* u, int nu, int* nev) {
  if (h == 0) {
    Timer t(c, 4);
    k_obs<<<std::min(c->dev.C, c->sms * 8), NTHR, 0, c->stream>>>(c->dev, mode, cand, outv, c->slice_w[0], c->slice_m[0],
                                                                 sampler_id, sweep, u, nu, nev);
    ++c->launches;
    CK(c, cudaGetLastError());
    return GPF_OK;
  }
  const int g = (h - 1) / 2, which = (h - 1) % 2;
  return launch_prior(c, g, mode ? 2 + which : which, cand, outv, sampler_id, sweep, u, nu, nev);
}

}  // namespace

extern "C" {

int gpf_version(void) { return 100; }

const char* gpf_build_info(void) {
  return "gpfanova_b200 sm_100a fp64 (mma.sync m8n8k4 f64 tile engine), built " __DATE__ " " __TIME__;
}

// ---------------------------------------------------------------------------------------------------------
// stateless operators
// ---------------------------------------------------------------------------------------------------------
int gpf_rbf_K(const double* d_x, int n, int p, const double* d_log10_sigma, const double* d_log10_ls, int batch,
              double* d_K, void* stream) {
  if (n < 1 || p < 1 || batch < 0 || !d_x || !d_K) { g_err = "gpf_rbf_K: bad argument"; return GPF_ERR_ARG; }
  if (batch == 0) return GPF_OK;
  const size_t smem = ((size_t)n * p + n) * sizeof(double);
  if (smem > smem_limit()) { g_err = "gpf_rbf_K: n*p too large for shared memory"; return GPF_ERR_UNSUPPORTED; }
  CKS(allow_smem(k_rbf_dense, smem));
  k_rbf_dense<<<batch, NTHR, smem, (cudaStream_t)stream>>>(d_x, n, p, d_log10_sigma, d_log10_ls, d_K);
  CKS(cudaGetLastError());
  return GPF_OK;
}

static int dense_factor_common(const double* d_A, int n, int batch, int maxtries, double add, double* d_out,
                               double* d_logdet, int* d_status, void* stream, bool inverse) {
  if (n < 1 || n > GPF_MAX_N || batch < 0 || !d_A || !d_out || !d_status) { g_err = "bad argument (1 <= n <= 512)"; return GPF_ERR_ARG; }
  if (batch == 0) return GPF_OK;
  const int NT = (n + TB - 1) / TB;
  const size_t smem = engine_smem_doubles(NT, 0) * sizeof(double);
  if (smem > smem_limit()) { g_err = "n too large for the tile engine"; return GPF_ERR_UNSUPPORTED; }
  int dev = 0, sms = 148;
  CKS(cudaGetDevice(&dev));
  CKS(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int grid = std::min(batch, 2 * sms);
  const size_t stride = (size_t)plt_tiles(NT) * TILE;
  double* ws = nullptr;
  CKS(cudaMalloc(&ws, (size_t)grid * stride * sizeof(double)));
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e;
  if (inverse) {
    e = allow_smem(k_kinv_dense, smem);
    if (e == cudaSuccess) {
      k_kinv_dense<<<grid, NTHR, smem, st>>>(d_A, n, NT, batch, maxtries, add, d_out, d_status, ws, stride);
      e = cudaGetLastError();
    }
  } else {
    e = allow_smem(k_jitchol_dense, smem);
    if (e == cudaSuccess) {
      k_jitchol_dense<<<grid, NTHR, smem, st>>>(d_A, n, NT, batch, maxtries, add, d_out, d_logdet, d_status, ws, stride);
      e = cudaGetLastError();
    }
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFree(ws);
  if (e != cudaSuccess) { g_err = cudaGetErrorString(e); return GPF_ERR_CUDA; }
  return GPF_OK;
}

int gpf_jitchol(const double* d_A, int n, int batch, int maxtries, double* d_L, double* d_logdet, int* d_status,
                void* stream) {
  return dense_factor_common(d_A, n, batch, maxtries, 0.0, d_L, d_logdet, d_status, stream, false);
}

int gpf_kinv(const double* d_A, int n, int batch, double offset, double* d_Ainv, int* d_status, void* stream) {
  return dense_factor_common(d_A, n, batch, 5, offset, d_Ainv, nullptr, d_status, stream, true);
}

int gpf_rbf_dist(const double* d_x, int n, int p, const double* d_log10_ls, int batch, double* d_out, void* stream) {
  if (n < 1 || p < 1 || batch < 0 || !d_x || !d_out || !d_log10_ls) { g_err = "gpf_rbf_dist: bad argument"; return GPF_ERR_ARG; }
  if (batch == 0) return GPF_OK;
  const size_t smem = ((size_t)n * p + n) * sizeof(double);
  if (smem > smem_limit()) { g_err = "gpf_rbf_dist: n*p too large for shared memory"; return GPF_ERR_UNSUPPORTED; }
  CKS(allow_smem(k_rbf_aux, smem));
  k_rbf_aux<<<batch, NTHR, smem, (cudaStream_t)stream>>>(d_x, n, p, nullptr, d_log10_ls, d_out, 1);
  CKS(cudaGetLastError());
  return GPF_OK;
}

int gpf_rbf_dK(const double* d_x, int n, int p, const double* d_log10_sigma, const double* d_log10_ls, int cross, int batch,
               double* d_out, void* stream) {
  if (n < 1 || p < 1 || batch < 0 || !d_x || !d_out || !d_log10_ls || !d_log10_sigma) { g_err = "gpf_rbf_dK: bad argument"; return GPF_ERR_ARG; }
  if (batch == 0) return GPF_OK;
  const size_t smem = ((size_t)n * p + n) * sizeof(double);
  if (smem > smem_limit()) { g_err = "gpf_rbf_dK: n*p too large for shared memory"; return GPF_ERR_UNSUPPORTED; }
  CKS(allow_smem(k_rbf_aux, smem));
  k_rbf_aux<<<batch, NTHR, smem, (cudaStream_t)stream>>>(d_x, n, p, d_log10_sigma, d_log10_ls, d_out, cross ? 3 : 2);
  CKS(cudaGetLastError());
  return GPF_OK;
}

static int gp_common(bool sample, const double* d_x, int n, int p, const double* d_ls, const double* d_ll, int batch, const double* d_F,
                     int nf, double add_rel, double add_abs, int maxtries, unsigned long long seed, unsigned stream_id, double* d_out,
                     double* d_logdet, int* d_status, void* stream) {
  if (n < 1 || n > GPF_MAX_N || p < 1 || batch < 0 || nf < 1 || !d_x || !d_ls || !d_ll || !d_out || !d_status || (!sample && !d_F)) {
    g_err = "gpf_gp_*: bad argument (1 <= n <= 512, nf >= 1, non-null arrays)";
    return GPF_ERR_ARG;
  }
  if (batch == 0) return GPF_OK;
  const int NT = (n + TB - 1) / TB, NE = sample ? 0 : (nf + TB - 1) / TB;
  size_t smem = (engine_smem_doubles(NT, NE) + (size_t)n * p + n) * sizeof(double);
  if (sample) smem += (size_t)16 * (NT * TB + 4) * sizeof(double);
  if (smem > smem_limit()) { g_err = "gpf_gp_*: n / number of functions exceed the shared-memory budget of the tile engine"; return GPF_ERR_UNSUPPORTED; }
  int dev = 0, sms = 148;
  CKS(cudaGetDevice(&dev));
  CKS(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int grid = std::min(batch, sms);
  GpArgs a{};
  a.x = d_x; a.n = n; a.p = p; a.NT = NT; a.batch = batch; a.nf = nf; a.maxtries = maxtries;
  a.lsig = d_ls; a.lls = d_ll; a.F = d_F; a.add_rel = add_rel; a.add_abs = add_abs; a.out = d_out; a.logdet = d_logdet;
  a.status = d_status;
  a.ws_stride = ((size_t)plt_tiles(NT) + (size_t)NE * NT) * TILE;
  a.k0 = (unsigned)seed; a.k1 = (unsigned)(seed >> 32); a.stream_id = stream_id;
  double* ws = nullptr;
  CKS(cudaMalloc(&ws, (size_t)grid * a.ws_stride * sizeof(double)));
  a.ws = ws;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = sample ? allow_smem(k_gp_sample, smem) : allow_smem(k_gp_loglik, smem);
  if (e == cudaSuccess) {
    if (sample) k_gp_sample<<<grid, NTHR, smem, st>>>(a);
    else k_gp_loglik<<<grid, NTHR, smem, st>>>(a);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFree(ws);
  if (e != cudaSuccess) { g_err = cudaGetErrorString(e); return GPF_ERR_CUDA; }
  return GPF_OK;
}

int gpf_gp_loglik(const double* d_x, int n, int p, const double* d_log10_sigma, const double* d_log10_ls, int batch, const double* d_F,
                  int nf, double add_rel, double add_abs, int maxtries, double* d_ll, double* d_logdet, int* d_status, void* stream) {
  return gp_common(false, d_x, n, p, d_log10_sigma, d_log10_ls, batch, d_F, nf, add_rel, add_abs, maxtries, 0ull, 0u, d_ll, d_logdet,
                   d_status, stream);
}

int gpf_gp_sample(const double* d_x, int n, int p, const double* d_log10_sigma, const double* d_log10_ls, int batch, int nsamp,
                  const double* d_z_inject, unsigned long long seed, unsigned stream_id, double add_rel, double add_abs, int maxtries,
                  double* d_out, int* d_status, void* stream) {
  return gp_common(true, d_x, n, p, d_log10_sigma, d_log10_ls, batch, d_z_inject, nsamp, add_rel, add_abs, maxtries, seed, stream_id,
                   d_out, nullptr, d_status, stream);
}

int gpf_function_derivative(const double* d_x, int n, int p, const double* d_log10_sigma, const double* d_log10_ls, int batch,
                            const double* d_F, int nf, const double* d_z_inject, unsigned long long seed, unsigned stream_id,
                            double offset, int maxtries, double* d_dF, double* d_mu, double* d_cov, int* d_status, void* stream) {
  if (n < 1 || n > GPF_MAX_N || p < 1 || batch < 0 || nf < 1 || nf > 32 || !d_x || !d_log10_sigma || !d_log10_ls || !d_F || !d_dF || !d_status) {
    g_err = "gpf_function_derivative: bad argument (1 <= n <= 512, 1 <= nf <= 32 functions per call)";
    return GPF_ERR_ARG;
  }
  if (batch == 0) return GPF_OK;
  const int NT = (n + TB - 1) / TB, NE = (nf + n + TB - 1) / TB;
  const size_t smem = (engine_smem_doubles(NT, NE) + (size_t)n * p + n) * sizeof(double);
  if (smem > smem_limit()) { g_err = "gpf_function_derivative: n too large for the shared-memory budget (n + nf rows ride along)"; return GPF_ERR_UNSUPPORTED; }
  int dev = 0, sms = 148;
  CKS(cudaGetDevice(&dev));
  CKS(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int grid = std::min(batch, sms);
  DerivArgs a{};
  a.x = d_x; a.n = n; a.p = p; a.NT = NT; a.NE = NE; a.batch = batch; a.nf = nf; a.maxtries = maxtries;
  a.lsig = d_log10_sigma; a.lls = d_log10_ls; a.F = d_F; a.z = d_z_inject; a.offset = offset;
  a.dF = d_dF; a.mu = d_mu; a.cov = d_cov; a.status = d_status;
  a.ws_stride = deriv_ws_doubles(NT, NE, nf, n);
  a.k0 = (unsigned)seed; a.k1 = (unsigned)(seed >> 32); a.stream_id = stream_id;
  double* ws = nullptr;
  CKS(cudaMalloc(&ws, (size_t)grid * a.ws_stride * sizeof(double)));
  a.ws = ws;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = allow_smem(k_function_derivative, smem);
  if (e == cudaSuccess) {
    k_function_derivative<<<grid, NTHR, smem, st>>>(a);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFree(ws);
  if (e != cudaSuccess) { g_err = cudaGetErrorString(e); return GPF_ERR_CUDA; }
  return GPF_OK;
}

int gpf_compose_y(const double* d_F, const double* d_X, const double* d_sd, int batch, int n, int m, int f, unsigned long long seed,
                  unsigned stream_id, double* d_y, void* stream) {
  if (!d_F || !d_X || !d_sd || !d_y || batch < 0 || n < 1 || m < 1 || f < 1) { g_err = "gpf_compose_y: bad argument"; return GPF_ERR_ARG; }
  if (batch == 0) return GPF_OK;
  const size_t total = (size_t)batch * n * m;
  const int grid = (int)std::min<size_t>((total + 255) / 256, 148 * 8);
  k_compose_y<<<grid, 256, 0, (cudaStream_t)stream>>>(d_F, d_X, d_sd, batch, n, m, f, d_y, (unsigned)seed, (unsigned)(seed >> 32), stream_id);
  CKS(cudaGetLastError());
  return GPF_OK;
}

int gpf_fanova_transform(const double* d_F, long long rows, int f, int n, int nblk, const int* h_in0, const int* h_nin, const int* h_nout,
                         const int* h_fpv, const double* h_coef, double* d_out, void* stream) {
  if (!d_F || !d_out || rows < 0 || f < 1 || n < 1 || nblk < 0 || (nblk && (!h_in0 || !h_nin || !h_nout || !h_fpv || !h_coef))) {
    g_err = "gpf_fanova_transform: bad argument";
    return GPF_ERR_ARG;
  }
  if (rows == 0 || nblk == 0) return GPF_OK;
  std::vector<int> out0(nblk), coff(nblk), cblk, cj0;
  int Q = 0;
  size_t nc = 0;
  for (int b = 0; b < nblk; ++b) {
    if (h_in0[b] < 0 || h_nin[b] < 1 || h_in0[b] + h_nin[b] > f || h_nout[b] < 1) { g_err = "gpf_fanova_transform: block out of range"; return GPF_ERR_ARG; }
    out0[b] = Q;
    coff[b] = (int)nc;
    for (int j0 = 0; j0 < h_nout[b]; j0 += 8) { cblk.push_back(b); cj0.push_back(j0); }
    Q += h_nout[b] + (h_fpv[b] ? 1 : 0);
    nc += (size_t)h_nout[b] * h_nin[b];
  }
  // description -> device (one allocation)
  const size_t ni = (size_t)6 * nblk + 2 * cblk.size();
  std::vector<int> hi(ni);
  int* q = hi.data();
  std::copy(h_in0, h_in0 + nblk, q);
  std::copy(h_nin, h_nin + nblk, q + nblk);
  std::copy(out0.begin(), out0.end(), q + 2 * nblk);
  std::copy(h_nout, h_nout + nblk, q + 3 * nblk);
  std::copy(coff.begin(), coff.end(), q + 4 * nblk);
  std::copy(h_fpv, h_fpv + nblk, q + 5 * nblk);
  std::copy(cblk.begin(), cblk.end(), q + 6 * nblk);
  std::copy(cj0.begin(), cj0.end(), q + 6 * nblk + cblk.size());
  int* di = nullptr;
  double* dc = nullptr;
  CKS(cudaMalloc(&di, ni * sizeof(int)));
  cudaError_t e = cudaMalloc(&dc, nc * sizeof(double));
  cudaStream_t st = (cudaStream_t)stream;
  if (e == cudaSuccess) e = cudaMemcpyAsync(di, hi.data(), ni * sizeof(int), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(dc, h_coef, nc * sizeof(double), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) {
    TransformDesc d{};
    d.nchunk = (int)cblk.size();
    d.in0 = di; d.nin = di + nblk; d.out0 = di + 2 * nblk; d.nout = di + 3 * nblk; d.coef_off = di + 4 * nblk; d.fpv = di + 5 * nblk;
    d.chunk_blk = di + 6 * nblk; d.chunk_j0 = di + 6 * nblk + cblk.size();
    d.coef = dc; d.f = f; d.n = n; d.Q = Q;
    for (long long r0 = 0; r0 < rows && e == cudaSuccess; r0 += 32768) {          // grid.x limit is generous; keep launches bounded
      const int nr = (int)std::min<long long>(32768, rows - r0);
      k_transform<<<dim3(nr, d.nchunk), 128, 0, st>>>(d_F + (size_t)r0 * f * n, d_out + (size_t)r0 * Q * n, d);
      k_transform_fpv<<<nr, 128, 0, st>>>(d_out + (size_t)r0 * Q * n, d, nblk);
      e = cudaGetLastError();
    }
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFree(di);
  if (dc) cudaFree(dc);
  if (e != cudaSuccess) { g_err = cudaGetErrorString(e); return GPF_ERR_CUDA; }
  return GPF_OK;
}

int gpf_chain_diagnostics(const double* d_hist, int S, int C, long long cols, int col0, int ncol, int split, double* d_rhat,
                          double* d_mean, double* d_var, void* stream) {
  if (!d_hist || !d_rhat || !d_mean || !d_var || S < 2 || C < 1 || col0 < 0 || ncol < 1 || col0 + ncol > cols) {
    g_err = "gpf_chain_diagnostics: bad argument (S >= 2 stored rows, columns inside the row)";
    return GPF_ERR_ARG;
  }
  const int halves = (split && S >= 4) ? 2 : 1, len = S / halves, nseq = halves * C;
  double* tmp = nullptr;
  CKS(cudaMalloc(&tmp, (size_t)2 * nseq * ncol * sizeof(double)));
  double* mean = tmp;
  double* var = tmp + (size_t)nseq * ncol;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaSuccess;
  for (int h = 0; h < halves && e == cudaSuccess; ++h) {
    // split chains: first and last `len` rows (the middle row of an odd S is dropped)
    const int s0 = (h == 0) ? 0 : S - len;
    for (int c0 = 0; c0 < C && e == cudaSuccess; c0 += 32768) {
      const int nc = std::min(32768, C - c0);
      k_seq_moments<<<dim3((ncol + 127) / 128, nc), 128, 0, st>>>(d_hist + (size_t)c0 * cols, S, C, (size_t)cols, col0, ncol, s0, s0 + len,
                                                               mean + ((size_t)h * C + c0) * ncol, var + ((size_t)h * C + c0) * ncol);
      e = cudaGetLastError();
    }
  }
  if (e == cudaSuccess) {
    k_rhat<<<(ncol + 127) / 128, 128, 0, st>>>(mean, var, nseq, ncol, len, d_rhat, d_mean, d_var);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFree(tmp);
  if (e != cudaSuccess) { g_err = cudaGetErrorString(e); return GPF_ERR_CUDA; }
  return GPF_OK;
}

int gpf_row_sort(double* d_v, int rows, int S, void* stream) {
  if (!d_v || rows < 0 || S < 1 || S > 8192) { g_err = "gpf_row_sort: 1 <= S <= 8192"; return GPF_ERR_ARG; }
  if (rows == 0) return GPF_OK;
  int P2 = 1;
  while (P2 < S) P2 <<= 1;
  CKS(allow_smem(k_row_sort, (size_t)P2 * sizeof(double)));
  k_row_sort<<<rows, 256, (size_t)P2 * sizeof(double), (cudaStream_t)stream>>>(d_v, S, P2);
  CKS(cudaGetLastError());
  return GPF_OK;
}

int gpf_function_interval(const double* d_samples, int batch, int S, int n, double alpha, double start, double tol, int maxiter,
                          double* d_mean, double* d_std, double* d_eps, int* d_iters, void* stream) {
  if (!d_samples || !d_mean || !d_std || !d_eps || !d_iters || batch < 0 || S < 1 || n < 1 || alpha < 0 || alpha > 1) {
    g_err = "gpf_function_interval: bad argument (must provide alpha between 0 and 1)";
    return GPF_ERR_ARG;
  }
  if (batch == 0) return GPF_OK;
  // interval.py:66-67: alpha moved to the closest of k/S, k = 0..S-1 (first one on ties)
  double best = 1e300, aeff = 0.0;
  for (int k = 0; k < S; ++k) {
    const double v = (double)k / (double)S, dlt = std::fabs(v - alpha);
    if (dlt < best) { best = dlt; aeff = v; }
  }
  const size_t smem = (size_t)2 * n * sizeof(double);
  CKS(allow_smem(k_band, smem));
  k_band<<<batch, 256, smem, (cudaStream_t)stream>>>(d_samples, S, n, aeff, alpha, start, tol, maxiter, d_mean, d_std, d_eps, d_iters);
  CKS(cudaGetLastError());
  return GPF_OK;
}

// ---------------------------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------------------------
int gpf_create(gpf_ctx** out, const gpf_model_desc* d) {
  if (!out || !d) { g_err = "gpf_create: null argument"; return GPF_ERR_ARG; }
  *out = nullptr;
  if (d->n < 1 || d->n > GPF_MAX_N || d->p < 1 || d->m < 1 || d->f < 1 || d->P < 1 || d->chains < 1 || !d->h_x || !d->h_y ||
      !d->h_X || !d->h_group_of_fn || !d->h_slice_w || !d->h_slice_m) {
    g_err = "gpf_create: bad model description (1 <= n <= 512, positive sizes, non-null arrays)";
    return GPF_ERR_ARG;
  }
  gpf_ctx* c = new gpf_ctx();
  auto fail = [&](int rc) { g_err = c->err; gpf_destroy(c); return rc; };
  c->device = d->device;
  {
    cudaError_t e = cudaSetDevice(d->device);
    if (e != cudaSuccess) { c->err = std::string("cudaSetDevice: ") + cudaGetErrorString(e); return fail(GPF_ERR_CUDA); }
    e = cudaDeviceGetAttribute(&c->sms, cudaDevAttrMultiProcessorCount, d->device);
    if (e != cudaSuccess) { c->err = cudaGetErrorString(e); return fail(GPF_ERR_CUDA); }
    e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    for (int i = 0; i < 4 && e == cudaSuccess; ++i) {
      e = cudaStreamCreateWithFlags(&c->aux[i], cudaStreamNonBlocking);
      if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_join[i], cudaEventDisableTiming);
      if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_user[i], cudaEventDisableTiming | cudaEventBlockingSync);
    }
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->lane2.st, cudaStreamNonBlocking);
    for (int i = 0; i < 4 && e == cudaSuccess; ++i) {
      e = cudaStreamCreateWithFlags(&c->lane2.aux[i], cudaStreamNonBlocking);
      if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->lane2.join[i], cudaEventDisableTiming);
    }
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->lane2.fork, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->lane2.begin, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->lane2.end, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->copy, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_copy, cudaEventDisableTiming);
    if (e != cudaSuccess) { c->err = cudaGetErrorString(e); return fail(GPF_ERR_CUDA); }
  }
  CtxDev& v = c->dev;
  v.n = d->n; v.p = d->p; v.m = d->m; v.f = d->f; v.P = d->P; v.C = d->chains; v.H = 1 + 2 * d->P;
  v.NT = (d->n + TB - 1) / TB;
  v.chain_offset = d->chain_offset;
  v.maxtries = 5;
  v.offset = d->offset; v.prior_jitter = d->prior_jitter; v.prior_lb = d->prior_lb; v.prior_ub = d->prior_ub;
  v.k0 = (unsigned)d->seed; v.k1 = (unsigned)(d->seed >> 32);
  c->group_of_fn.assign(d->h_group_of_fn, d->h_group_of_fn + d->f);
  c->groups.assign(d->P, {});
  for (int i = 0; i < d->f; ++i) {
    const int g = c->group_of_fn[i];
    if (g < 0 || g >= d->P) { c->err = "gpf_create: group_of_fn out of range"; return fail(GPF_ERR_ARG); }
    c->groups[g].push_back(i);
  }
  c->slice_w.assign(d->h_slice_w, d->h_slice_w + v.H);
  c->slice_m.assign(d->h_slice_m, d->h_slice_m + v.H);
  // sampler list in the reference's order (base.py:57-84)
  c->order.push_back({0, 0});
  int ngmax = 1;
  for (int g = 0; g < d->P; ++g) {
    if (c->groups[g].empty()) { c->err = "gpf_create: empty prior group"; return fail(GPF_ERR_ARG); }
    for (int fn : c->groups[g]) c->order.push_back({1, fn});
    c->order.push_back({0, 1 + 2 * g});
    c->order.push_back({0, 2 + 2 * g});
    ngmax = std::max(ngmax, (int)c->groups[g].size());
  }
  v.NEmax = (ngmax + TB - 1) / TB;
  c->lb_valid.assign(d->P, 0);
  // uniform grid?  then every RBF matrix is Toeplitz: K[i][j] = k(|i-j| h) and needs one exp per lag
  std::vector<double> d2(d->n, 0.0);
  c->toeplitz = false;
  if (d->p == 1) {
    const long double x0 = d->h_x[0], h = d->n > 1 ? ((long double)d->h_x[d->n - 1] - x0) / (d->n - 1) : 0.0L;
    const double tol = 1e-14 * (std::fabs(d->h_x[0]) + std::fabs(d->h_x[d->n - 1]));
    bool uni = true;
    for (int i = 0; i < d->n && uni; ++i) uni = std::fabs((double)(x0 + i * h - (long double)d->h_x[i])) <= tol;
    if (uni) {
      c->toeplitz = true;
      for (int k = 0; k < d->n; ++k) d2[k] = (double)((k * h) * (k * h));
    }
  }
  // symmetric Toeplitz => centrosymmetric: everything n^3 runs on the two half-size blocks (n even; small n gains nothing)
  c->centro_ok = c->toeplitz && d->n % 2 == 0 && d->n >= 64;
  v.centro = c->centro_ok ? 1 : 0;
  // X^T X (decides whether the function draws of a sweep are independent) and the classes of functions that share their
  // conditional precision
  std::vector<double> G((size_t)d->f * d->f, 0.0);
  for (int a = 0; a < d->f; ++a)
    for (int b = 0; b < d->f; ++b) {
      double s = 0.0;
      for (int o = 0; o < d->m; ++o) s += d->h_X[(size_t)o * d->f + a] * d->h_X[(size_t)o * d->f + b];
      G[(size_t)a * d->f + b] = s;
    }
  bool nomiss = true;
  for (size_t i = 0; i < (size_t)d->n * d->m; ++i) nomiss = nomiss && (d->h_y[i] == d->h_y[i]);
  bool diag = true;
  for (int a = 0; a < d->f; ++a) {
    // a function no observation loads on (all-zero design column): rank deficient (base.py:42-44 logs it and draws from
    // the prior); the least-squares shortcuts divide by (X^T X)[a,a], so such a design takes the general path
    if (!(G[(size_t)a * d->f + a] > 0.0)) diag = false;
    for (int b = 0; b < d->f; ++b)
      // couplings below 1e-13 sqrt(G_aa G_bb) are rounding noise of the contrast construction (helmertConvert
      // rescales the Helmert columns with square roots) and do not make the draws dependent
      if (a != b && std::fabs(G[(size_t)a * d->f + b]) > 1e-13 * std::sqrt(G[(size_t)a * d->f + a] * G[(size_t)b * d->f + b])) diag = false;
  }
  c->indep = diag && nomiss;
  // sampler ids of the functions, classes per group (same (X^T X)[f,f] within a group)
  std::vector<int> fl, coff(1, 0), cfn, gcoff(1, 0), coff1(1, 0), gcoff1(1, 0);
  std::vector<unsigned> sl, csid;
  std::vector<unsigned> sid_of(d->f, 0u);
  for (size_t i = 0; i < c->order.size(); ++i)
    if (c->order[i].kind == 1) { fl.push_back(c->order[i].arg); sl.push_back((unsigned)i); sid_of[c->order[i].arg] = (unsigned)i; }
  int maxcls = 1;
  for (int g = 0; g < d->P; ++g) {
    const std::vector<int>& grp = c->groups[g];
    std::vector<char> used(grp.size(), 0);
    for (size_t i = 0; i < grp.size(); ++i) {
      if (used[i]) continue;
      const double gi = G[(size_t)grp[i] * d->f + grp[i]];
      int cnt = 0;
      for (size_t j = i; j < grp.size(); ++j) {
        if (used[j]) continue;
        if (std::fabs(G[(size_t)grp[j] * d->f + grp[j]] - gi) > 1e-13 * gi) continue;
        used[j] = 1; cfn.push_back(grp[j]); csid.push_back(sid_of[grp[j]]); ++cnt;
        coff1.push_back((int)cfn.size());
      }
      coff.push_back((int)cfn.size());
      maxcls = std::max(maxcls, cnt);
    }
    gcoff.push_back((int)coff.size() - 1);
    gcoff1.push_back((int)coff1.size() - 1);
  }
  c->ncls = (int)coff.size() - 1;
  (void)maxcls;
  {
    size_t gmax = 0;
    for (auto& grp : c->groups) gmax = std::max(gmax, grp.size());
    c->schurfn_ok = c->toeplitz && c->indep && d->n <= 256 && d->n >= 8 && gmax <= (size_t)SFN_NG;
    // the Schur function step uses the full Cholesky factor of K_o as the square root of the prior draw; the dense kernels
    // then do the same (no centrosymmetric split), so that every path of one context implements one draw contract
    if (c->schurfn_ok) v.centro = 0;
  }
  // shared-memory budgets
  const size_t xsd = (size_t)d->n * d->p + d->n;
  c->smem_kinv = (engine_smem_doubles(v.NT, 0) + xsd) * sizeof(double);
  c->smem_kinv2 = (2 * engine_smem_doubles((d->n / 2 + TB - 1) / TB, 0) + xsd) * sizeof(double);   // centrosymmetric split
  c->smem_prior.resize(d->P);
  size_t smax = std::max(c->smem_kinv, c->smem_kinv2);
  for (int mode = 0; mode < 3; ++mode) {
    if (mode == 2 && !c->centro_ok) continue;
    if (mode == 1 && !c->toeplitz) continue;
    const int nsub = mode == 2 ? 2 : 1, NTm = ((mode == 2 ? d->n / 2 : d->n) + TB - 1) / TB;
    c->smem_fn[mode] = function_smem_doubles(d->n, d->p, v.NT, mode) * sizeof(double);
    c->smem_fact[mode] = ((size_t)nsub * engine_smem_doubles(NTm, 0) + xsd) * sizeof(double);
    c->smem_fused[mode] = fused_smem_doubles(d->n, d->p, mode) * sizeof(double);
    smax = std::max(smax, std::max(c->smem_fn[mode], std::max(c->smem_fact[mode], c->indep ? c->smem_fused[mode] : (size_t)0)));
  }
  for (int g = 0; g < d->P; ++g) {
    const int NE = ((int)c->groups[g].size() + TB - 1) / TB;
    c->smem_prior[g] = (engine_smem_doubles(v.NT, NE) + xsd) * sizeof(double);
    smax = std::max(smax, c->smem_prior[g]);
  }
  c->schur_warps.assign(d->P, 0);
  for (int g = 0; g < d->P; ++g) {
    const size_t per_warp = schur_warp_doubles(d->n, (int)c->groups[g].size()) * sizeof(double);
    const int w = (int)std::min<size_t>(NW, (110 * 1024) / per_warp);   // two CTAs per SM
    c->schur_warps[g] = w;                                            // 0: vectors do not fit -> dense engine
  }
  if (smax > smem_limit()) {
    c->err = "gpf_create: n / prior-group size exceed the shared-memory budget of the tile engine";
    return fail(GPF_ERR_UNSUPPORTED);
  }
  {
    const size_t sp = *std::max_element(c->smem_prior.begin(), c->smem_prior.end());
    cudaError_t e = allow_smem(k_group_kinv<0>, c->smem_kinv);
    if (e == cudaSuccess) e = allow_smem(k_group_kinv<1>, c->smem_kinv);
    if (e == cudaSuccess) e = allow_smem(k_group_kinv<2>, c->smem_kinv2);
    if (e == cudaSuccess) e = allow_smem(k_function<0>, c->smem_fn[0]);
    if (e == cudaSuccess) e = allow_smem(k_function<1>, c->smem_fn[1]);
    if (e == cudaSuccess) e = allow_smem(k_function<2>, c->smem_fn[2]);
    if (e == cudaSuccess) e = allow_smem(k_group_factor<0>, c->smem_fact[0]);
    if (e == cudaSuccess) e = allow_smem(k_group_factor<1>, c->smem_fact[1]);
    if (e == cudaSuccess) e = allow_smem(k_group_factor<2>, c->smem_fact[2]);
    if (e == cudaSuccess) e = allow_smem(k_group_fused<0>, c->smem_fused[0]);
    if (e == cudaSuccess) e = allow_smem(k_group_fused<1>, c->smem_fused[1]);
    if (e == cudaSuccess) e = allow_smem(k_group_fused<2>, c->smem_fused[2]);
    if (e == cudaSuccess) e = allow_smem(k_prior<true>, sp);
    if (e == cudaSuccess) e = allow_smem(k_prior<false>, sp);
    if (e == cudaSuccess) e = schur_set_attributes();
    if (e == cudaSuccess) e = sweepk_set_attributes();
    if (e != cudaSuccess) { c->err = std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e); return fail(GPF_ERR_CUDA); }
  }
  c->grid = std::min(v.C, 2 * c->sms);
  const int ws_ctas = 2 * c->sms;     // batched launches (several functions / groups per launch) may fill the GPU even for small C
  // device buffers
  const size_t tstride = (size_t)plt_tiles(v.NT) * TILE;
  const int NTh = (d->n / 2 + TB - 1) / TB;
  const size_t hstride2 = 2 * (size_t)plt_tiles(NTh) * TILE;                      // the two half-size factors
  v.lb_stride = c->centro_ok ? std::max(tstride, hstride2) : tstride;
  // per-CTA workspace: dense slice kernel / general function kernel (matrix + right-hand-side tile rows), fused kernel
  // (factor(s) of K_o, factor(s) of C, one right-hand-side tile row each)
  v.ws_stride = tstride + (size_t)std::max(v.NEmax, 1) * v.NT * TILE;
  v.ws_stride = std::max(v.ws_stride, 2 * tstride + (size_t)v.NT * TILE);
  if (c->centro_ok) v.ws_stride = std::max(v.ws_stride, 2 * hstride2 + 2 * (size_t)NTh * TILE);
  v.fws_stride = std::max(fused_scratch_doubles(d->n, 1), c->centro_ok ? fused_scratch_doubles(d->n, 2) : (size_t)0);
  double *dx, *dyT, *dX, *dF, *dh, *dLB, *dBj, *dws, *dd2, *dG, *dZ, *dfws;
  unsigned char* dfull;
  int *dgof, *dgf, *dgo, *dst, *dsched;
  unsigned long long* dcn;
  int rc;
  if ((rc = dev_alloc(c, &dx, (size_t)d->n * d->p)) || (rc = dev_alloc(c, &dyT, (size_t)d->n * d->m)) ||
      (rc = dev_alloc(c, &dX, (size_t)d->m * d->f)) || (rc = dev_alloc(c, &dF, (size_t)v.C * d->f * d->n)) ||
      (rc = dev_alloc(c, &dh, (size_t)v.C * v.H)) || (rc = dev_alloc(c, &dLB, (size_t)d->P * v.C * v.lb_stride)) ||
      (rc = dev_alloc(c, &dBj, (size_t)d->P * v.C)) ||
      (rc = dev_alloc(c, &dws, (size_t)ws_ctas * v.ws_stride)) || (rc = dev_alloc(c, &dgof, (size_t)d->f)) ||
      (rc = dev_alloc(c, &dgf, (size_t)d->f)) || (rc = dev_alloc(c, &dgo, (size_t)d->P + 1)) ||
      (rc = dev_alloc(c, &dst, (size_t)v.C)) || (rc = dev_alloc(c, &dcn, (size_t)CNT_N)) ||
      (rc = dev_alloc(c, &dsched, (size_t)16)) || (rc = dev_alloc(c, &dd2, (size_t)d->n)) || (rc = dev_alloc(c, &dG, (size_t)d->f * d->f)) ||
      (rc = dev_alloc(c, &dZ, (size_t)d->f * d->n)) || (rc = dev_alloc(c, &dfull, (size_t)d->n)) ||
      (rc = dev_alloc(c, &dfws, (size_t)ws_ctas * v.fws_stride)))
    return fail(rc);
  v.fws = dfws;
  // X^T y (observed entries) and the per-time-point "nothing missing" flag for the residual projection
  std::vector<double> Z((size_t)d->f * d->n, 0.0);
  std::vector<unsigned char> full(d->n, 1);
  for (int t = 0; t < d->n; ++t)
    for (int o = 0; o < d->m; ++o) {
      const double yv = d->h_y[(size_t)t * d->m + o];
      if (yv != yv) { full[t] = 0; continue; }
      for (int a = 0; a < d->f; ++a) Z[(size_t)a * d->n + t] += d->h_X[(size_t)o * d->f + a] * yv;
    }
  std::vector<double> yT((size_t)d->n * d->m);
  for (int t = 0; t < d->n; ++t)
    for (int o = 0; o < d->m; ++o) yT[(size_t)o * d->n + t] = d->h_y[(size_t)t * d->m + o];
  std::vector<int> gf, go(1, 0);
  for (int g = 0; g < d->P; ++g) {
    gf.insert(gf.end(), c->groups[g].begin(), c->groups[g].end());
    go.push_back((int)gf.size());
  }
  {
    v.bhat = nullptr;
    v.rss0 = 0.0;
    if (c->indep) {   // least-squares functions: observation likelihood about the fit (k_obs), right-hand sides of the fused kernel
      std::vector<double> bh((size_t)d->f * d->n);
      for (int a = 0; a < d->f; ++a)
        for (int t = 0; t < d->n; ++t) bh[(size_t)a * d->n + t] = Z[(size_t)a * d->n + t] / G[(size_t)a * d->f + a];
      long double rss = 0.0L;
      for (int t = 0; t < d->n; ++t)
        for (int o = 0; o < d->m; ++o) {
          long double pred = 0.0L;
          for (int a = 0; a < d->f; ++a) pred += (long double)d->h_X[(size_t)o * d->f + a] * bh[(size_t)a * d->n + t];
          const long double r = (long double)d->h_y[(size_t)t * d->m + o] - pred;
          rss += r * r;
        }
      double* dbh;
      if ((rc = dev_alloc(c, &dbh, bh.size()))) return fail(rc);
      cudaMemcpy(dbh, bh.data(), sizeof(double) * bh.size(), cudaMemcpyHostToDevice);
      v.bhat = dbh;
      c->d_bhat = dbh;
      v.rss0 = (double)rss;
    }
    // groups ordered by descending work for the fused kernel's dynamic hand-out (largest items first)
    std::vector<int> gorder(d->P);
    for (int g = 0; g < d->P; ++g) gorder[g] = g;
    std::stable_sort(gorder.begin(), gorder.end(), [&](int a, int b) { return c->groups[a].size() > c->groups[b].size(); });
    if ((rc = dev_alloc(c, &c->d_fn_all, fl.size())) || (rc = dev_alloc(c, &c->d_sid_all, sl.size())) ||
        (rc = dev_alloc(c, &c->d_cls_off, coff.size())) || (rc = dev_alloc(c, &c->d_cls_fn, cfn.size())) ||
        (rc = dev_alloc(c, &c->d_cls_sid, csid.size())) || (rc = dev_alloc(c, &c->d_gcls_off, gcoff.size())) ||
        (rc = dev_alloc(c, &c->d_cls1_off, coff1.size())) || (rc = dev_alloc(c, &c->d_gcls1_off, gcoff1.size())) ||
        (rc = dev_alloc(c, &c->d_gorder, gorder.size())))
      return fail(rc);
    cudaMemcpy(c->d_fn_all, fl.data(), sizeof(int) * fl.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(c->d_sid_all, sl.data(), sizeof(unsigned) * sl.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(c->d_cls_off, coff.data(), sizeof(int) * coff.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(c->d_cls_fn, cfn.data(), sizeof(int) * cfn.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(c->d_cls_sid, csid.data(), sizeof(unsigned) * csid.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(c->d_gcls_off, gcoff.data(), sizeof(int) * gcoff.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(c->d_cls1_off, coff1.data(), sizeof(int) * coff1.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(c->d_gcls1_off, gcoff1.data(), sizeof(int) * gcoff1.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(c->d_gorder, gorder.data(), sizeof(int) * gorder.size(), cudaMemcpyHostToDevice);
    c->h_gorder = gorder;
  }
  cudaError_t e = cudaMemcpy(dx, d->h_x, sizeof(double) * d->n * d->p, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(dyT, yT.data(), sizeof(double) * yT.size(), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(dX, d->h_X, sizeof(double) * d->m * d->f, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(dgof, c->group_of_fn.data(), sizeof(int) * d->f, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(dgf, gf.data(), sizeof(int) * gf.size(), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(dgo, go.data(), sizeof(int) * go.size(), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(dd2, d2.data(), sizeof(double) * d->n, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(dG, G.data(), sizeof(double) * G.size(), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(dZ, Z.data(), sizeof(double) * Z.size(), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(dfull, full.data(), full.size(), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemset(dF, 0, sizeof(double) * v.C * d->f * d->n);
  if (e == cudaSuccess) e = cudaMemset(dh, 0, sizeof(double) * v.C * v.H);
  if (e == cudaSuccess) e = cudaMemset(dst, 0, sizeof(int) * v.C);
  if (e == cudaSuccess) e = cudaMemset(dcn, 0, sizeof(unsigned long long) * CNT_N);
  if (e != cudaSuccess) { c->err = std::string("gpf_create: upload: ") + cudaGetErrorString(e); return fail(GPF_ERR_CUDA); }
  v.x = dx; v.yT = dyT; v.X = dX; v.F = dF; v.hyper = dh; v.Kinv = nullptr; v.LB = dLB; v.Bjit = dBj; v.ws = dws;
  v.sched = dsched;
  v.d2 = dd2; v.G = dG; v.Z = dZ; v.full = dfull;
  v.group_of_fn = dgof; v.group_fns = dgf; v.group_off = dgo; v.status = dst; v.counters = dcn;
  *out = c;
  return GPF_OK;
}

void gpf_destroy(gpf_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  collect_profile(c);
  for (auto e : c->ev_pool) cudaEventDestroy(e);
  for (void* p : c->allocs) cudaFree(p);
  for (int i = 0; i < 4; ++i) {
    if (c->aux[i]) { cudaStreamSynchronize(c->aux[i]); cudaStreamDestroy(c->aux[i]); }
    if (c->ev_join[i]) cudaEventDestroy(c->ev_join[i]);
    if (c->ev_user[i]) cudaEventDestroy(c->ev_user[i]);
  }
  if (c->ev_fork) cudaEventDestroy(c->ev_fork);
  for (int i = 0; i < 4; ++i) {
    if (c->lane2.aux[i]) { cudaStreamSynchronize(c->lane2.aux[i]); cudaStreamDestroy(c->lane2.aux[i]); }
    if (c->lane2.join[i]) cudaEventDestroy(c->lane2.join[i]);
  }
  if (c->lane2.st) { cudaStreamSynchronize(c->lane2.st); cudaStreamDestroy(c->lane2.st); }
  if (c->lane2.fork) cudaEventDestroy(c->lane2.fork);
  if (c->lane2.begin) cudaEventDestroy(c->lane2.begin);
  if (c->lane2.end) cudaEventDestroy(c->lane2.end);
  if (c->copy) { cudaStreamSynchronize(c->copy); cudaStreamDestroy(c->copy); }
  if (c->ev_copy) cudaEventDestroy(c->ev_copy);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

const char* gpf_last_error(gpf_ctx* c) { return c ? c->err.c_str() : g_err.c_str(); }
void* gpf_stream(gpf_ctx* c) { return c ? (void*)c->stream : nullptr; }

int gpf_sync(gpf_ctx* c) {
  CK(c, cudaSetDevice(c->device));
  CK(c, cudaStreamSynchronize(c->stream));
  CK(c, cudaStreamSynchronize(c->copy));
  return GPF_OK;
}

int gpf_set_state(gpf_ctx* c, const double* h_F, const double* h_hyper) {
  CK(c, cudaSetDevice(c->device));
  const CtxDev& v = c->dev;
  if (h_F) CK(c, cudaMemcpyAsync(v.F, h_F, sizeof(double) * v.C * v.f * v.n, cudaMemcpyHostToDevice, c->stream));
  if (h_hyper) {
    CK(c, cudaMemcpyAsync(v.hyper, h_hyper, sizeof(double) * v.C * v.H, cudaMemcpyHostToDevice, c->stream));
    std::fill(c->lb_valid.begin(), c->lb_valid.end(), 0);
  }
  CK(c, cudaStreamSynchronize(c->stream));
  return GPF_OK;
}

int gpf_get_state(gpf_ctx* c, double* h_F, double* h_hyper) {
  CK(c, cudaSetDevice(c->device));
  const CtxDev& v = c->dev;
  if (h_F) CK(c, cudaMemcpyAsync(h_F, v.F, sizeof(double) * v.C * v.f * v.n, cudaMemcpyDeviceToHost, c->stream));
  if (h_hyper) CK(c, cudaMemcpyAsync(h_hyper, v.hyper, sizeof(double) * v.C * v.H, cudaMemcpyDeviceToHost, c->stream));
  CK(c, cudaStreamSynchronize(c->stream));
  return GPF_OK;
}

int gpf_memcpy_d2h(gpf_ctx* c, void* h_dst, const void* d_src, size_t bytes) {
  CK(c, cudaSetDevice(c->device));
  CK(c, cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, c->stream));
  return GPF_OK;
}

int gpf_memcpy_d2h_event(gpf_ctx* c, void* h_dst, const void* d_src, size_t bytes, int idx) {
  if (idx < 0 || idx > 3) { c->err = "gpf_memcpy_d2h_event: idx must be 0..3"; return GPF_ERR_ARG; }
  CK(c, cudaSetDevice(c->device));
  CK(c, cudaEventRecord(c->ev_copy, c->stream));                 // everything queued so far on the ctx stream ...
  CK(c, cudaStreamWaitEvent(c->copy, c->ev_copy, 0));            // ... precedes the copy, which runs on its own stream
  CK(c, cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, c->copy));
  CK(c, cudaEventRecord(c->ev_user[idx], c->copy));
  return GPF_OK;
}

int gpf_record_event(gpf_ctx* c, int idx) {
  if (idx < 0 || idx > 3) { c->err = "gpf_record_event: idx must be 0..3"; return GPF_ERR_ARG; }
  CK(c, cudaSetDevice(c->device));
  CK(c, cudaEventRecord(c->ev_user[idx], c->stream));
  return GPF_OK;
}

int gpf_wait_event(gpf_ctx* c, int idx) {
  if (idx < 0 || idx > 3) { c->err = "gpf_wait_event: idx must be 0..3"; return GPF_ERR_ARG; }
  CK(c, cudaSetDevice(c->device));
  CK(c, cudaEventSynchronize(c->ev_user[idx]));
  return GPF_OK;
}

double* gpf_device_F(gpf_ctx* c) { return c->dev.F; }
double* gpf_device_hyper(gpf_ctx* c) { return c->dev.hyper; }

int gpf_invalidate(gpf_ctx* c) {   // call after writing gpf_device_hyper() directly
  std::fill(c->lb_valid.begin(), c->lb_valid.end(), 0);
  return GPF_OK;
}

int gpf_get_status(gpf_ctx* c, int* h_status, int clear) {
  CK(c, cudaSetDevice(c->device));
  std::vector<int> tmp(c->dev.C);
  CK(c, cudaMemcpyAsync(tmp.data(), c->dev.status, sizeof(int) * c->dev.C, cudaMemcpyDeviceToHost, c->stream));
  if (clear) CK(c, cudaMemsetAsync(c->dev.status, 0, sizeof(int) * c->dev.C, c->stream));
  CK(c, cudaStreamSynchronize(c->stream));
  if (c->d_swk) {
    int e = 0;
    CK(c, cudaMemcpy(&e, c->d_swk + 2 * c->dev.C, sizeof(int), cudaMemcpyDeviceToHost));
    if (e) { c->err = "k_chain_sweeps: a unit waited too long for its chain's y_sigma step (internal error)"; return GPF_ERR_CUDA; }
  }
  bool bad = false;
  for (int i = 0; i < c->dev.C; ++i) {
    if (h_status) h_status[i] = tmp[i];
    if (tmp[i] < 0) bad = true;
  }
  if (bad) { c->err = "not positive definite, even with jitter."; return GPF_ERR_NOT_PD; }
  return GPF_OK;
}

int gpf_get_counters(gpf_ctx* c, unsigned long long* h, int clear) {
  CK(c, cudaSetDevice(c->device));
  unsigned long long tmp[CNT_N];
  CK(c, cudaMemcpyAsync(tmp, c->dev.counters, sizeof tmp, cudaMemcpyDeviceToHost, c->stream));
  if (clear) CK(c, cudaMemsetAsync(c->dev.counters, 0, sizeof tmp, c->stream));
  CK(c, cudaStreamSynchronize(c->stream));
  for (int i = 0; i < 8; ++i) h[i] = tmp[i];
  return GPF_OK;
}

int gpf_group_kinv(gpf_ctx* c, int g, double* d_out) {
  CK(c, cudaSetDevice(c->device));
  const CtxDev& v = c->dev;
  if (g < -1 || g >= v.P || (g < 0 && d_out)) { c->err = "gpf_group_kinv: bad group"; return GPF_ERR_ARG; }
  // the factor of K_g + offset I that the function steps use (the sweep itself never forms the inverse) ...
  int rc = (g < 0) ? launch_factor(c, 0, v.P) : launch_factor(c, g);
  if (rc) return rc;
  // ... and, on request, the inverse itself (Kernel.K_inv as an operator on the chains' current hyper-parameters)
  if (d_out) {
    rc = launch_kinv(c, g);
    if (rc) return rc;
    const size_t tstride = (size_t)plt_tiles(v.NT) * TILE;
    k_tiles_to_dense<<<std::min(v.C, 1024), NTHR, 0, c->stream>>>(c->d_kinv_tmp, tstride, v.n, v.NT, v.C, d_out, 1);
    ++c->launches;
    CK(c, cudaGetLastError());
  }
  return GPF_OK;
}

int gpf_function_step(gpf_ctx* c, int fn, unsigned sampler_id, unsigned sweep, const double* d_z_inject, double* d_mu_out,
                      double* d_LC_out, double* d_mt_out, double* d_nt_out) {
  CK(c, cudaSetDevice(c->device));
  if (fn < 0 || fn >= c->dev.f) { c->err = "gpf_function_step: bad function index"; return GPF_ERR_ARG; }
  return launch_function(c, fn, sampler_id, sweep, d_z_inject, FnOut{d_mu_out, d_LC_out, d_mt_out, d_nt_out});
}

int gpf_obs_loglik(gpf_ctx* c, const double* d_cand, double* d_out) {
  CK(c, cudaSetDevice(c->device));
  if (!d_cand || !d_out) { c->err = "gpf_obs_loglik: null argument"; return GPF_ERR_ARG; }
  return launch_slice(c, 0, 0, d_cand, d_out, 0, 0, nullptr, 0, nullptr);
}

int gpf_prior_loglik(gpf_ctx* c, int g, int which, const double* d_cand, double* d_out) {
  CK(c, cudaSetDevice(c->device));
  if (g < 0 || g >= c->dev.P || which < 0 || which > 1 || !d_cand || !d_out) { c->err = "gpf_prior_loglik: bad argument"; return GPF_ERR_ARG; }
  return launch_slice(c, 1 + 2 * g + which, 0, d_cand, d_out, 0, 0, nullptr, 0, nullptr);
}

int gpf_slice_step(gpf_ctx* c, int h, unsigned sampler_id, unsigned sweep, const double* d_u_inject, int n_inject,
                   int* d_nevals_out) {
  CK(c, cudaSetDevice(c->device));
  if (h < 0 || h >= c->dev.H) { c->err = "gpf_slice_step: bad hyper-parameter index"; return GPF_ERR_ARG; }
  return launch_slice(c, h, 1, nullptr, nullptr, sampler_id, sweep, d_u_inject, d_u_inject ? n_inject : 0, d_nevals_out);
}

int gpf_n_samplers(gpf_ctx* c) { return (int)c->order.size(); }

int gpf_sweep(gpf_ctx* c, int n_sweeps, int thin, unsigned sweep0, const int* h_order, double* d_history,
              long long history_capacity_rows, long long* rows_out) {
  CK(c, cudaSetDevice(c->device));
  const CtxDev& v = c->dev;
  const int ns = (int)c->order.size();
  const size_t cols = (size_t)v.H + (size_t)v.f * v.n;
  long long rows = 0;
  if (h_order)
    for (long long i = 0; i < (long long)n_sweeps * ns; ++i)
      if (h_order[i] < 0 || h_order[i] >= ns) { c->err = "gpf_sweep: sampler id out of range in h_order"; return GPF_ERR_ARG; }
  cudaEvent_t t0 = nullptr, t1 = nullptr;
  if (c->profile) {
    for (double& m : c->ms) m = 0.0;
    cudaEventCreate(&t0); cudaEventCreate(&t1);
    cudaEventRecord(t0, c->stream);
  }
  bool one_launch = !h_order && sweepk_active(c);
  if (one_launch) {
    int rc = launch_chain_sweeps(c, n_sweeps, thin, sweep0, d_history, history_capacity_rows, &rows);
    if (rc) return rc;
  } else if (!h_order && pipe_active(c)) {
    int rc = pipelined_sweeps(c, n_sweeps, thin, sweep0, d_history, history_capacity_rows, &rows);
    if (rc) return rc;
    one_launch = true;          // (nothing left for the per-stage loop below)
  }
  for (int s = 0; s < n_sweeps && !one_launch; ++s) {
    const unsigned sweep = sweep0 + (unsigned)s;
    if (!h_order && c->indep && c->batch) {
      // Balanced design, nothing missing: given the hyper-parameters the function draws of a sweep do not depend on
      // each other (X^T X diagonal), and a group's slice steps only read that group's functions, so the reference's
      // order  y_sigma, [functions, sigma, lengthscale] per group  is reproduced exactly by
      //   y_sigma | K_inv of all groups | all functions | slices per group   (same Philox counters, same inputs)
      int rc = launch_slice(c, 0, 1, nullptr, nullptr, 0u, sweep, nullptr, 0, nullptr);
      if (!rc) rc = launch_all_functions(c, sweep);
      // the groups' slice kernels are independent of each other: run them concurrently on auxiliary streams (a small
      // chain count does not fill the GPU alone; a large one leaves a tail of long slice steps that the next group's
      // CTAs fill) -- unless per-kernel timing is on: the timers sit on the main stream
      bool conc = c->profile != 1;
      for (int g = 0; g < c->dev.P; ++g) conc = conc && uses_schur(c, g);   // the dense slice kernel owns the per-CTA workspace
      if (conc && !rc) CK(c, cudaEventRecord(c->ev_fork, c->stream));
      int used = 0;
      // largest groups first: their slice steps are the longest, and the smaller groups' CTAs then fill their tail
      std::vector<std::pair<int, int>> todo;     // (sampler index, group)
      for (int k = 0; k < ns; ++k)
        if (c->order[k].kind == 0 && c->order[k].arg >= 1 && (c->order[k].arg - 1) % 2 == 0) todo.push_back({k, (c->order[k].arg - 1) / 2});
      if (conc)
        std::stable_sort(todo.begin(), todo.end(), [&](const std::pair<int, int>& a, const std::pair<int, int>& b) {
          return c->groups[a.second].size() > c->groups[b.second].size();
        });
      std::vector<char> forked(4, 0);
      for (size_t q = 0; q < todo.size() && !rc; ++q) {
        const int k = todo[q].first, g = todo[q].second, slot = conc ? g % 4 : -1;
        if (conc && !forked[slot]) { CK(c, cudaStreamWaitEvent(c->aux[slot], c->ev_fork, 0)); forked[slot] = 1; ++used; }
        rc = launch_prior(c, g, 4, nullptr, nullptr, (unsigned)k, sweep, nullptr, 0, nullptr, slot);
      }
      if (conc && !rc)
        for (int i = 0; i < 4; ++i)
          if (forked[i]) {
            CK(c, cudaEventRecord(c->ev_join[i], c->aux[i]));
            CK(c, cudaStreamWaitEvent(c->stream, c->ev_join[i], 0));
          }
      (void)used;
      if (rc) return rc;
    } else
    for (int k = 0; k < ns; ++k) {
      const int sid = h_order ? h_order[(size_t)s * ns + k] : k;
      const Samp& sp = c->order[sid];
      int rc;
      if (sp.kind == 1) rc = launch_function(c, sp.arg, (unsigned)sid, sweep, nullptr, FnOut{nullptr, nullptr, nullptr, nullptr});
      else {
        // prior{g}_sigma immediately followed by prior{g}_lengthscale (the reference's fixed order): one fused launch
        const int nxt = (k + 1 < ns) ? (h_order ? h_order[(size_t)s * ns + k + 1] : k + 1) : -1;
        if (sp.arg >= 1 && (sp.arg - 1) % 2 == 0 && nxt == sid + 1 && c->order[nxt].kind == 0 && c->order[nxt].arg == sp.arg + 1) {
          rc = launch_prior(c, (sp.arg - 1) / 2, 4, nullptr, nullptr, (unsigned)sid, sweep, nullptr, 0, nullptr);
          ++k;
        } else rc = launch_slice(c, sp.arg, 1, nullptr, nullptr, (unsigned)sid, sweep, nullptr, 0, nullptr);
      }
      if (rc) return rc;
    }
    if (d_history && (thin <= 0 || (s + 1) % thin == 0)) {
      if (rows >= history_capacity_rows) { c->err = "gpf_sweep: history buffer full"; return GPF_ERR_ARG; }
      Timer t(c, 5);
      double* row = d_history + (size_t)rows * v.C * cols;
      CK(c, cudaMemcpy2DAsync(row, cols * sizeof(double), v.hyper, v.H * sizeof(double), v.H * sizeof(double), v.C,
                              cudaMemcpyDeviceToDevice, c->stream));
      CK(c, cudaMemcpy2DAsync(row + v.H, cols * sizeof(double), v.F, (size_t)v.f * v.n * sizeof(double),
                              (size_t)v.f * v.n * sizeof(double), v.C, cudaMemcpyDeviceToDevice, c->stream));
      ++rows;
    }
  }
  if (rows_out) *rows_out = rows;
  if (c->profile) {
    cudaEventRecord(t1, c->stream);
    cudaEventSynchronize(t1);
    float t = 0.f;
    cudaEventElapsedTime(&t, t0, t1);
    collect_profile(c);
    c->ms[0] = t;
    cudaEventDestroy(t0); cudaEventDestroy(t1);
  }
  return GPF_OK;
}

int gpf_set_option(gpf_ctx* c, int option, int value) {
  if (option == 0) { c->use_schur = value; return GPF_OK; }
  if (option == 1) { c->batch = value; return GPF_OK; }
  if (option == 2) { c->share = value; return GPF_OK; }
  if (option == 3) {   // centrosymmetric split of K_inv on uniform grids (on by default where it applies)
    c->dev.centro = (value && c->centro_ok) ? 1 : 0;
    for (auto& q : c->lb_valid) q = 0;
    return GPF_OK;
  }
  if (option == 4) { c->dev.bhat = value ? c->d_bhat : nullptr; return GPF_OK; }
  if (option == 5) { c->use_schurfn = value; return GPF_OK; }
  if (option == 7) { c->use_sweepk = value; return GPF_OK; }
  if (option == 8) { c->pipelines = value; return GPF_OK; }
  if (option == 9) { c->cohorts = value < 1 ? 1 : (value > 2 ? 2 : value); return GPF_OK; }
  c->err = "gpf_set_option: unknown option";
  return GPF_ERR_ARG;
}

// effective setting of an option (what the kernels of the next sweep will do); -1 for unknown options
int gpf_get_option(gpf_ctx* c, int option) {
  switch (option) {
    case 0: return c->use_schur;
    case 1: return c->batch;
    case 2: return c->share;
    case 3: return c->dev.centro;
    case 4: return c->dev.bhat != nullptr;
    case 5: return schurfn_active(c) ? 1 : 0;
    case 6: return c->indep ? 1 : 0;       // (read-only) balanced design, nothing missing
    case 7: return sweepk_active(c) ? 1 : 0;
    case 8: return pipe_active(c) ? c->pipelines : 0;
    case 9: return c->cohorts;
    default: return -1;
  }
}

int gpf_set_profile(gpf_ctx* c, int on) { c->profile = on; return GPF_OK; }

// phase cycle counters of the fused function kernel (builds with -DGPF_FTRACE only; zeros otherwise): read and clear
int gpf_debug_trace(gpf_ctx* c, unsigned long long* h24) {
  CK(c, cudaSetDevice(c->device));
  for (int i = 0; i < 24; ++i) h24[i] = 0;
#ifdef GPF_FTRACE
  CK(c, cudaStreamSynchronize(c->stream));
  CK(c, cudaMemcpyFromSymbol(h24, gpf::gpf_ftrace, 8 * sizeof(unsigned long long)));
  unsigned long long z[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  CK(c, cudaMemcpyToSymbol(gpf::gpf_ftrace, z, 8 * sizeof(unsigned long long)));
#endif
  return GPF_OK;
}

int gpf_get_profile(gpf_ctx* c, double* h_ms6, unsigned long long* h_launches) {
  if (h_ms6) for (int i = 0; i < 6; ++i) h_ms6[i] = c->ms[i];
  if (h_launches) *h_launches = c->launches;
  return GPF_OK;
}

}  // extern "C"
