// Shared-factor local Kriging: the job-index kernel and the dispatcher over the instantiation units of
// krige_local_shr.cuh.
#include <algorithm>
#include <cstdlib>

#include "gsm_internal.cuh"
#include "krige_local_jobs.cuh"

int gsm_shr_entry_d0_b1(gsm_ctx*, const VarioDev&, const ModelDev&, int, int, const TargetsDev&, const JobIdx*, long long, int, const int*, double*, double*, unsigned char*);
int gsm_shr_entry_d0_b2(gsm_ctx*, const VarioDev&, const ModelDev&, int, int, const TargetsDev&, const JobIdx*, long long, int, const int*, double*, double*, unsigned char*);
int gsm_shr_entry_d0_b4(gsm_ctx*, const VarioDev&, const ModelDev&, int, int, const TargetsDev&, const JobIdx*, long long, int, const int*, double*, double*, unsigned char*);
int gsm_shr_entry_d1_b1(gsm_ctx*, const VarioDev&, const ModelDev&, int, int, const TargetsDev&, const JobIdx*, long long, int, const int*, double*, double*, unsigned char*);
int gsm_shr_entry_d1_b2(gsm_ctx*, const VarioDev&, const ModelDev&, int, int, const TargetsDev&, const JobIdx*, long long, int, const int*, double*, double*, unsigned char*);
int gsm_shr_entry_d1_b4(gsm_ctx*, const VarioDev&, const ModelDev&, int, int, const TargetsDev&, const JobIdx*, long long, int, const int*, double*, double*, unsigned char*);

namespace {

constexpr int IX_WARPS = 8;
constexpr int IX_HSIZE = 512;
constexpr int IX_UMAX = 256;   // 8 lists x 32 entries

struct IxWarp {
  int hkey[IX_HSIZE];                 // record position, -1: empty
  int hval[IX_HSIZE];                 // (multiplicity << 8) | unique id
  int upos[GSM_JOB_PMAX], uh[GSM_JOB_PMAX], urank[GSM_JOB_PMAX];
  unsigned skey[GSM_JOB_PMAX];
};

// One warp per job: pool the 8 neighbour lists of a patch and write its descriptor (krige_local_jobs.cuh).
__global__ void __launch_bounds__(IX_WARPS * 32)
job_index_kernel(TargetsDev tg, GroupMap gm, int k, int kmax, int smax, int minn, const int* __restrict__ pos,
                 const int* __restrict__ cnt, JobIdx* __restrict__ out) {
  __shared__ IxWarp Wsh[IX_WARPS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  IxWarp& W = Wsh[warp];
  for (int e = lane; e < IX_HSIZE; e += 32) {
    W.hkey[e] = -1;
    W.hval[e] = 0;
  }
  __syncwarp();
  const long long grp = blockIdx.x * (long long)IX_WARPS + warp;
  if (grp >= gm.ngroups) return;
  JobIdx& J = out[grp];

  // members of the patch: lanes 0..7
  long long tt = -1;
  int n = 0, fl = 0;
  int ci[3] = {0, 0, 0};
  if (lane < 8) {
    tt = group_member(gm, tg, grp, lane, ci);
    if (tt >= 0) {
      n = cnt[tt];
      fl = 1;
      if (n < minn || n <= 0) {
        n = 0;
        fl |= 2;
      }
      n = min(n, min(k, kmax));
    }
  }
  const int nactive = __popc(__ballot_sync(0xffffffffu, n > 0));
  // the eight lists, first hash probe of every entry in flight together
  int key8[8], old8[8], hs8[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const long long tt_i = __shfl_sync(0xffffffffu, tt, i);
    const int n_i = __shfl_sync(0xffffffffu, n, i);
    key8[i] = (lane < n_i) ? __ldg(pos + tt_i * (long long)k + lane) : -1;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    old8[i] = 0;
    if (key8[i] >= 0) old8[i] = atomicCAS(&W.hkey[(int)(((unsigned)key8[i] * 0x9E3779B1u) >> 23)], -1, key8[i]);
  }
  const unsigned lt_mask = (1u << lane) - 1u;
  int P = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int key = key8[i];
    bool owner = false;
    int h = 0;
    if (key >= 0) {
      h = (int)(((unsigned)key * 0x9E3779B1u) >> 23);
      int old = old8[i];
      while (old != -1 && old != key) {
        h = (h + 1) & (IX_HSIZE - 1);
        old = atomicCAS(&W.hkey[h], -1, key);
      }
      owner = old == -1;
    }
    const unsigned ob = __ballot_sync(0xffffffffu, owner);
    int add = key >= 0 ? 256 : 0;   // multiplicity
    if (owner) {
      const int uid = P + __popc(ob & lt_mask);   // < 256
      add += uid;
      if (uid < GSM_JOB_PMAX) {
        W.upos[uid] = key;
        W.uh[uid] = h;
      }
    }
    if (key >= 0) atomicAdd(&W.hval[h], add);
    P += __popc(ob);
    hs8[i] = h;
  }
  __syncwarp();

  if (lane < 8) {
    J.n[lane] = (unsigned char)n;
    J.fl[lane] = (unsigned char)fl;
    J.tt[lane] = (int)tt;
    J.ci[lane][0] = ci[0];
    J.ci[lane][1] = ci[1];
    J.ci[lane][2] = ci[2];
  }
  if (P > GSM_JOB_PMAX) {   // the union does not fit the pool: the solve kernel assembles this job directly
    if (lane == 0) {
      J.P = -1;
      J.S = 0;
    }
    return;
  }
  // rank the pool: samples common to every active member first, then ascending record position
  unsigned mykey[2];
  int c = 0;
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const int u = lane + 32 * q;
    bool common = false;
    mykey[q] = 0xffffffffu;
    if (u < P) {
      common = (W.hval[W.uh[u]] >> 8) == nactive;
      mykey[q] = (common ? 0u : 0x80000000u) | (unsigned)W.upos[u];
      W.skey[u] = mykey[q];
    }
    c += __popc(__ballot_sync(0xffffffffu, common));
  }
  __syncwarp();
  int rank[2] = {0, 0};
#pragma unroll 4
  for (int w = 0; w < P; ++w) {
    const unsigned kw = W.skey[w];
    rank[0] += (kw < mykey[0]) ? 1 : 0;
    rank[1] += (kw < mykey[1]) ? 1 : 0;
  }
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const int u = lane + 32 * q;
    if (u < P) {
      W.urank[u] = rank[q];
      J.upos[rank[q]] = (int)(mykey[q] & 0x7fffffffu);
    }
  }
  __syncwarp();
  if (lane == 0) {
    J.P = P;
    J.S = nactive > 0 ? min(smax, c >> 3) : 0;
  }
  // slot order of every member: ascending pool rank, i.e. an entry of rank r sits in slot popc(mask below r)
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    unsigned lo = 0u, hi = 0u;
    int r = -1;
    if (key8[i] >= 0) {
      r = W.urank[W.hval[hs8[i]] & 255];
      if (r < 32) lo = 1u << r;
      else hi = 1u << (r - 32);
    }
    lo = __reduce_or_sync(0xffffffffu, lo);
    hi = __reduce_or_sync(0xffffffffu, hi);
    const int nn = __popc(lo) + __popc(hi);
    if (r >= 0) {
      const int slot = r < 32 ? __popc(lo & ((1u << r) - 1u)) : __popc(lo) + __popc(hi & ((1u << (r - 32)) - 1u));
      J.sid[i][slot] = (unsigned char)r;
    }
    if (lane >= nn) J.sid[i][lane] = 255;
  }
}

}  // namespace

// Pool the neighbour lists of every patch of the launch: one 768-byte descriptor per job (krige_local_jobs.cuh).
int gsm_build_jobs(gsm_ctx* ctx, const TargetsDev& tg, int k, int nbc, int smax, int minneighbors, const int* d_pos,
                   const int* d_cnt, const JobIdx** out_jobs, long long* out_ngroups) {
  const GroupMap gm = gsm_make_group_map(tg);
  GSM_CHECK(gsm_reserve(ctx, ctx->job_idx, (size_t)gm.ngroups * sizeof(JobIdx)));
  JobIdx* jobs = (JobIdx*)ctx->job_idx.p;
  const unsigned ixblocks = (unsigned)((gm.ngroups + IX_WARPS - 1) / IX_WARPS);
  auto mark = [&]() {
    cudaEvent_t e;
    if (cudaEventCreate(&e) == cudaSuccess) {
      cudaEventRecord(e, ctx->stream);
      ctx->idx_marks.push_back(e);
    }
  };
  mark();
  job_index_kernel<<<ixblocks, IX_WARPS * 32, 0, ctx->stream>>>(tg, gm, k, 8 * nbc, smax, minneighbors, d_pos, d_cnt, jobs);
  ctx->tim.launches++;
  mark();
  GSM_CUDA(ctx, cudaGetLastError());
  *out_jobs = jobs;
  *out_ngroups = gm.ngroups;
  return GSM_OK;
}

int gsm_launch_local_krige_shr(gsm_ctx* ctx, const VarioDev& v, const ModelDev& m, int ncon, int centered,
                               const TargetsDev& tg, int k, int minneighbors, const int* d_pos, const int* d_cnt,
                               double* d_mean, double* d_var, unsigned char* d_status) {
  if (k > 32 || k < 1) return GSM_ERR_UNSUPPORTED;
  if (tg.count <= 0) return GSM_OK;
  if (tg.count >= (1LL << 31)) return GSM_ERR_UNSUPPORTED;   // descriptors hold 32-bit relative indices
  switch (ncon) {   // (instantiated constraint counts, krige_local_shr.cuh)
    case 0: case 1: case 2: case 3: case 4: case 6: case 10: break;
    default: return GSM_ERR_UNSUPPORTED;
  }
  const int nbc = k <= 8 ? 1 : (k <= 16 ? 2 : 4);
  const bool d3 = ctx->dim >= 3;
  const JobIdx* jobs = nullptr;
  long long ngroups = 0;
  const int smax = std::min(nbc - 1, ctx->opt_shr_smax);   // developer option "shr_smax": cap the sharing
  GSM_CHECK(gsm_build_jobs(ctx, tg, k, nbc, smax, minneighbors, d_pos, d_cnt, &jobs, &ngroups));
#define GSM_SHR_CALL(D, B) gsm_shr_entry_d##D##_b##B(ctx, v, m, ncon, centered, tg, jobs, ngroups, k, d_pos, d_mean, d_var, d_status)
  if (!d3) return nbc == 1 ? GSM_SHR_CALL(0, 1) : (nbc == 2 ? GSM_SHR_CALL(0, 2) : GSM_SHR_CALL(0, 4));
  return nbc == 1 ? GSM_SHR_CALL(1, 1) : (nbc == 2 ? GSM_SHR_CALL(1, 2) : GSM_SHR_CALL(1, 4));
#undef GSM_SHR_CALL
}
