// Jobs of the shared-factor local-Kriging kernel (krige_local_shr.cuh): a job is a compact patch of 8 targets
// (4 x 2 in 2-D, 2 x 2 x 2 in 3-D, 8 consecutive targets otherwise).  The job-index kernel (krige_local_shr.cu)
// turns the 8 neighbour lists of a patch (models.jl:162-184: one search per target) into
//   * the pool of distinct samples of the patch, RANKED: samples common to every active member first, then
//     ascending record position -- the canonical slot order of every system of the patch,
//   * the slot order of every member (pool ranks ascending) and the integer grid coordinates of its target,
//   * S, the number of 8-wide block columns all members share.
#pragma once
#include "gsm_internal.cuh"

constexpr int GSM_JOB_PMAX = 64;   // pool capacity; a larger union is solved by direct assembly (P = -1)

struct alignas(128) JobIdx {
  unsigned char sid[8][32];        // pool rank of every slot of every member, canonical order (255: empty slot)
  int P;                           // distinct samples (0..64); -1: the union does not fit the pool
  int S;                           // shared leading block columns (8 samples each)
  unsigned char n[8];              // neighbours per member (0: missing / none)
  unsigned char fl[8];             // bit 0: write a result, bit 1: missing (fewer than minneighbors)
  int tt[8];                       // relative target index of each member (-1: none)
  int ci[8][3];                    // grid index of each member's target (grid targets only)
  int pad[26];
  int upos[GSM_JOB_PMAX];          // record position of every pooled sample, by rank
};
static_assert(sizeof(JobIdx) == 768, "JobIdx layout");
constexpr int GSM_JOB_WORDS = sizeof(JobIdx) / 4;   // 192 = 6 warps x 32 lanes

// Patch enumeration of one launch (a contiguous slab [first, first + count) of the target domain).
struct GroupMap {
  int sx, sy, sz;              // patch shape, sx * sy * sz == 8; sx == 8: eight consecutive targets
  long long nx, ny, nz;        // grid extent
  long long nxb, nyb;          // patches along x and y inside the covered range
  long long yb0, zb0;          // first patch row / plane of the covered range
  long long ngroups;
  int small;                   // every index fits 31 bits
};

#ifdef __CUDACC__
// relative target index (0 .. count-1) of member m of patch grp, or -1
__device__ __forceinline__ long long group_member(const GroupMap& gm, const TargetsDev& tg, long long grp, int m,
                                                  int* ci = nullptr) {
  if (ci) ci[0] = ci[1] = ci[2] = 0;
  if (grp >= gm.ngroups) return -1;
  if (gm.sx == 8) {
    const long long tt = grp * 8 + m;
    return tt < tg.count ? tt : -1;
  }
  const int i = m % gm.sx, j = (m / gm.sx) % gm.sy, l = m / (gm.sx * gm.sy);
  long long x, y, z;
  if (gm.small) {   // 32-bit index arithmetic (64-bit division costs hundreds of instructions)
    const unsigned g32 = (unsigned)grp, nxb = (unsigned)gm.nxb, nyb = (unsigned)gm.nyb;
    const unsigned r = g32 / nxb, xb = g32 - r * nxb, zq = r / nyb, yq = r - zq * nyb;
    x = xb * (unsigned)gm.sx + i;
    y = (yq + (unsigned)gm.yb0) * (unsigned)gm.sy + j;
    z = (zq + (unsigned)gm.zb0) * (unsigned)gm.sz + l;
  } else {
    const long long xb = grp % gm.nxb, r = grp / gm.nxb;
    x = xb * gm.sx + i;
    y = (r % gm.nyb + gm.yb0) * gm.sy + j;
    z = (r / gm.nyb + gm.zb0) * gm.sz + l;
  }
  if (x >= gm.nx || y >= gm.ny || z >= gm.nz) return -1;
  if (ci) {
    ci[0] = (int)x;
    ci[1] = (int)y;
    ci[2] = (int)z;
  }
  const long long tt = x + gm.nx * (y + gm.ny * z) - tg.first;
  return (tt >= 0 && tt < tg.count) ? tt : -1;
}
#endif

inline GroupMap gsm_make_group_map(const TargetsDev& tg) {
  GroupMap gm{};
  gm.nx = gm.ny = gm.nz = 1;
  gm.small = 1;
  const bool grid = tg.kind == GSM_TARGETS_GRID && tg.dim >= 2 && tg.dims[1] > 1 && tg.count > 0;
  if (!grid) {
    gm.sx = 8;
    gm.sy = gm.sz = 1;
    gm.nxb = gm.nyb = 1;
    gm.ngroups = (tg.count + 7) / 8;
    return gm;
  }
  gm.nx = tg.dims[0];
  gm.ny = tg.dims[1];
  gm.nz = tg.dim > 2 ? tg.dims[2] : 1;
  const long long r_lo = tg.first / gm.nx, r_hi = (tg.first + tg.count - 1) / gm.nx;
  const long long z_lo = r_lo / gm.ny, z_hi = r_hi / gm.ny;
  const long long y_lo = r_lo % gm.ny, y_hi = r_hi % gm.ny;
  if (gm.nz > 1 && z_hi > z_lo) {
    gm.sx = gm.sy = gm.sz = 2;
  } else {
    gm.sx = 4;
    gm.sy = 2;
    gm.sz = 1;
  }
  gm.zb0 = z_lo / gm.sz;
  const long long zb1 = z_hi / gm.sz;
  long long yb1;
  if (z_lo == z_hi) {
    gm.yb0 = y_lo / gm.sy;
    yb1 = y_hi / gm.sy;
  } else {
    gm.yb0 = 0;
    yb1 = (gm.ny - 1) / gm.sy;
  }
  gm.nyb = yb1 - gm.yb0 + 1;
  gm.nxb = (gm.nx + gm.sx - 1) / gm.sx;
  gm.ngroups = gm.nxb * gm.nyb * (zb1 - gm.zb0 + 1);
  gm.small = (gm.ngroups < (1LL << 31) && gm.nx + 8 < (1LL << 30) && gm.ny + 8 < (1LL << 30) &&
              gm.nz + 8 < (1LL << 30)) ? 1 : 0;
  return gm;
}
