// Yee FDTD half-steps, guard-cell copy, guard->interior current reduce, J/rho reset.
//
// Reference behaviour restated (not copied): core/maxwell/cpu.py:9-35,83-112 (numba, no FMA contraction),
// core/patch/sync_fields3d.c:84-620, sync_fields2d.c:43-255, core/current/cpu3d.c:185-240.
// These kernels are bit-exact against the reference: every fp64 operation is an explicit round-to-nearest
// intrinsic in the reference's association order, so nvcc cannot contract a*b+c into an FMA.
//
// HBM roofline note: all four kernels are pure streaming; one thread owns one (patch, cell).
// Threads run along z (the contiguous axis) so a warp reads 128-256 B runs; neighbours in y/x come from L1/L2.
#include <math.h>
#include <stdlib.h>
#include <vector>
#include "lpic_common.cuh"

namespace {

struct CellIdx {
    int p, i, j, k;
};

// t = global interior-cell number; all arithmetic after the first (block-uniform) division is 32-bit
__device__ __forceinline__ bool interior_cell(const Geom &g, i64 t, CellIdx &c) {
    const int per = g.nx * g.ny * g.nz;
    if (t >= (i64)per * g.npatch) return false;
    const i64 t0 = t - threadIdx.x;  // block-uniform: one 64-bit division per warp instead of a divergent one per lane
    const int p0 = (int)(t0 / per);
    int r = (int)(t0 - (i64)p0 * per) + (int)threadIdx.x;
    c.p = p0;
    if (r >= per) { c.p += r / per; r %= per; }
    c.k = r % g.nz;
    r /= g.nz;
    c.j = r % g.ny;
    c.i = r / g.ny;
    return true;
}

__device__ __forceinline__ int sidx(const Geom &g, int i, int j, int k) {
    return wrapneg(k, g.NZ) + g.NZ * (wrapneg(j, g.NY) + g.NY * wrapneg(i, g.NX));
}

// E += (dt c^2) curl B - (dt/eps0) J on interior cells; backward differences reach index -1 (low guard).
template <int DIM>
__global__ void __launch_bounds__(256) k_update_efield(Geom g, double *__restrict__ F, double bfactor, double jfactor,
                                                       const u8 *__restrict__ is_pml, const double *__restrict__ kappa, int nmax) {
    CellIdx c;
    if (!interior_cell(g, (i64)blockIdx.x * blockDim.x + threadIdx.x, c)) return;
    const size_t stride = (size_t)g.npatch * g.ncell;
    double *base = F + (size_t)c.p * g.ncell;
    double *ex = base + LPIC_EX * stride, *ey = base + LPIC_EY * stride, *ez = base + LPIC_EZ * stride;
    const double *bx = base + LPIC_BX * stride, *by = base + LPIC_BY * stride, *bz = base + LPIC_BZ * stride;
    const double *jx = base + LPIC_JX * stride, *jy = base + LPIC_JY * stride, *jz = base + LPIC_JZ * stride;
    const int o = sidx(g, c.i, c.j, c.k), xm = sidx(g, c.i - 1, c.j, c.k), ym = sidx(g, c.i, c.j - 1, c.k);
    const double bxc = bx[o], byc = by[o], bzc = bz[o];
    if (is_pml && is_pml[c.p]) {
        // kappa-scaled update of a CPML edge patch (core/boundary/cpml.py:342-362 2D, :437-457 3D); note that the 2D and
        // 3D forms associate the products differently -- both are reproduced literally
        const double *kp = kappa + (size_t)c.p * 6 * nmax;  // [e|b][axis][nmax], e first
        const double bfx = __ddiv_rn(bfactor, kp[c.i]), bfy = __ddiv_rn(bfactor, kp[nmax + c.j]);
        if (DIM == 3) {
            const int zm = sidx(g, c.i, c.j, c.k - 1);
            const double bfz = __ddiv_rn(bfactor, kp[2 * nmax + c.k]);
            const double cx = __dsub_rn(__ddiv_rn(__dmul_rn(bfy, __dsub_rn(bzc, bz[ym])), g.dy), __ddiv_rn(__dmul_rn(bfz, __dsub_rn(byc, by[zm])), g.dz));
            const double cy = __dsub_rn(__ddiv_rn(__dmul_rn(bfz, __dsub_rn(bxc, bx[zm])), g.dz), __ddiv_rn(__dmul_rn(bfx, __dsub_rn(bzc, bz[xm])), g.dx));
            const double cz = __dsub_rn(__ddiv_rn(__dmul_rn(bfx, __dsub_rn(byc, by[xm])), g.dx), __ddiv_rn(__dmul_rn(bfy, __dsub_rn(bxc, bx[ym])), g.dy));
            ex[o] = __dadd_rn(ex[o], __dsub_rn(cx, __dmul_rn(jfactor, jx[o])));
            ey[o] = __dadd_rn(ey[o], __dsub_rn(cy, __dmul_rn(jfactor, jy[o])));
            ez[o] = __dadd_rn(ez[o], __dsub_rn(cz, __dmul_rn(jfactor, jz[o])));
        } else {
            const double cx = __dmul_rn(bfy, __ddiv_rn(__dsub_rn(bzc, bz[ym]), g.dy));
            const double cy = __dmul_rn(bfx, __ddiv_rn(-__dsub_rn(bzc, bz[xm]), g.dx));
            const double cz = __dsub_rn(__dmul_rn(bfx, __ddiv_rn(__dsub_rn(byc, by[xm]), g.dx)), __dmul_rn(bfy, __ddiv_rn(__dsub_rn(bxc, bx[ym]), g.dy)));
            ex[o] = __dadd_rn(ex[o], __dsub_rn(cx, __dmul_rn(jfactor, jx[o])));
            ey[o] = __dadd_rn(ey[o], __dsub_rn(cy, __dmul_rn(jfactor, jy[o])));
            ez[o] = __dadd_rn(ez[o], __dsub_rn(cz, __dmul_rn(jfactor, jz[o])));
        }
        return;
    }
    if (DIM == 3) {
        const int zm = sidx(g, c.i, c.j, c.k - 1);
        // ex += bfactor*((bz[c]-bz[ym])/dy - (by[c]-by[zm])/dz) - jfactor*jx[c]       (cpu.py:92-97)
        double cx = __dsub_rn(__ddiv_rn(__dsub_rn(bzc, bz[ym]), g.dy), __ddiv_rn(__dsub_rn(byc, by[zm]), g.dz));
        double cy = __dsub_rn(__ddiv_rn(__dsub_rn(bxc, bx[zm]), g.dz), __ddiv_rn(__dsub_rn(bzc, bz[xm]), g.dx));
        double cz = __dsub_rn(__ddiv_rn(__dsub_rn(byc, by[xm]), g.dx), __ddiv_rn(__dsub_rn(bxc, bx[ym]), g.dy));
        ex[o] = __dadd_rn(ex[o], __dsub_rn(__dmul_rn(bfactor, cx), __dmul_rn(jfactor, jx[o])));
        ey[o] = __dadd_rn(ey[o], __dsub_rn(__dmul_rn(bfactor, cy), __dmul_rn(jfactor, jy[o])));
        ez[o] = __dadd_rn(ez[o], __dsub_rn(__dmul_rn(bfactor, cz), __dmul_rn(jfactor, jz[o])));
    } else {
        // 2D (cpu.py:22-24): ex += bfactor*((bz-bz[ym])/dy) ; ey += bfactor*(-(bz-bz[xm])/dx) ; ez as 3D
        double cx = __ddiv_rn(__dsub_rn(bzc, bz[ym]), g.dy);
        double cy = __ddiv_rn(-__dsub_rn(bzc, bz[xm]), g.dx);
        double cz = __dsub_rn(__ddiv_rn(__dsub_rn(byc, by[xm]), g.dx), __ddiv_rn(__dsub_rn(bxc, bx[ym]), g.dy));
        ex[o] = __dadd_rn(ex[o], __dsub_rn(__dmul_rn(bfactor, cx), __dmul_rn(jfactor, jx[o])));
        ey[o] = __dadd_rn(ey[o], __dsub_rn(__dmul_rn(bfactor, cy), __dmul_rn(jfactor, jy[o])));
        ez[o] = __dadd_rn(ez[o], __dsub_rn(__dmul_rn(bfactor, cz), __dmul_rn(jfactor, jz[o])));
    }
}

// B -= dt curl E on interior cells; forward differences reach index n (high guard).
template <int DIM>
__global__ void __launch_bounds__(256) k_update_bfield(Geom g, double *__restrict__ F, double dt, const u8 *__restrict__ is_pml,
                                                       const double *__restrict__ kappa, int nmax) {
    CellIdx c;
    if (!interior_cell(g, (i64)blockIdx.x * blockDim.x + threadIdx.x, c)) return;
    const size_t stride = (size_t)g.npatch * g.ncell;
    double *base = F + (size_t)c.p * g.ncell;
    const double *ex = base + LPIC_EX * stride, *ey = base + LPIC_EY * stride, *ez = base + LPIC_EZ * stride;
    double *bx = base + LPIC_BX * stride, *by = base + LPIC_BY * stride, *bz = base + LPIC_BZ * stride;
    const int o = sidx(g, c.i, c.j, c.k), xp = sidx(g, c.i + 1, c.j, c.k), yp = sidx(g, c.i, c.j + 1, c.k);
    const double exc = ex[o], eyc = ey[o], ezc = ez[o];
    if (is_pml && is_pml[c.p]) {  // core/boundary/cpml.py:364-377 (2D), :459-477 (3D)
        const double *kp = kappa + ((size_t)c.p * 6 + 3) * nmax;
        const double efx = __ddiv_rn(dt, kp[c.i]), efy = __ddiv_rn(dt, kp[nmax + c.j]);
        if (DIM == 3) {
            const int zp = sidx(g, c.i, c.j, c.k + 1);
            const double efz = __ddiv_rn(dt, kp[2 * nmax + c.k]);
            const double cx = __dsub_rn(__ddiv_rn(__dmul_rn(efy, __dsub_rn(ez[yp], ezc)), g.dy), __ddiv_rn(__dmul_rn(efz, __dsub_rn(ey[zp], eyc)), g.dz));
            const double cy = __dsub_rn(__ddiv_rn(__dmul_rn(efz, __dsub_rn(ex[zp], exc)), g.dz), __ddiv_rn(__dmul_rn(efx, __dsub_rn(ez[xp], ezc)), g.dx));
            const double cz = __dsub_rn(__ddiv_rn(__dmul_rn(efx, __dsub_rn(ey[xp], eyc)), g.dx), __ddiv_rn(__dmul_rn(efy, __dsub_rn(ex[yp], exc)), g.dy));
            bx[o] = __dsub_rn(bx[o], cx);
            by[o] = __dsub_rn(by[o], cy);
            bz[o] = __dsub_rn(bz[o], cz);
        } else {
            const double cx = __dmul_rn(efy, __ddiv_rn(__dsub_rn(ez[yp], ezc), g.dy));
            const double cy = __dmul_rn(efx, __ddiv_rn(-__dsub_rn(ez[xp], ezc), g.dx));
            const double cz = __dsub_rn(__dmul_rn(efx, __ddiv_rn(__dsub_rn(ey[xp], eyc), g.dx)), __dmul_rn(efy, __ddiv_rn(__dsub_rn(ex[yp], exc), g.dy)));
            bx[o] = __dsub_rn(bx[o], cx);
            by[o] = __dsub_rn(by[o], cy);
            bz[o] = __dsub_rn(bz[o], cz);
        }
        return;
    }
    if (DIM == 3) {
        const int zp = sidx(g, c.i, c.j, c.k + 1);
        double cx = __dsub_rn(__ddiv_rn(__dsub_rn(ez[yp], ezc), g.dy), __ddiv_rn(__dsub_rn(ey[zp], eyc), g.dz));
        double cy = __dsub_rn(__ddiv_rn(__dsub_rn(ex[zp], exc), g.dz), __ddiv_rn(__dsub_rn(ez[xp], ezc), g.dx));
        double cz = __dsub_rn(__ddiv_rn(__dsub_rn(ey[xp], eyc), g.dx), __ddiv_rn(__dsub_rn(ex[yp], exc), g.dy));
        bx[o] = __dsub_rn(bx[o], __dmul_rn(dt, cx));
        by[o] = __dsub_rn(by[o], __dmul_rn(dt, cy));
        bz[o] = __dsub_rn(bz[o], __dmul_rn(dt, cz));
    } else {
        double cx = __ddiv_rn(__dsub_rn(ez[yp], ezc), g.dy);
        double cy = __ddiv_rn(-__dsub_rn(ez[xp], ezc), g.dx);
        double cz = __dsub_rn(__ddiv_rn(__dsub_rn(ey[xp], eyc), g.dx), __ddiv_rn(__dsub_rn(ex[yp], exc), g.dy));
        bx[o] = __dsub_rn(bx[o], __dmul_rn(dt, cx));
        by[o] = __dsub_rn(by[o], __dmul_rn(dt, cy));
        bz[o] = __dsub_rn(bz[o], __dmul_rn(dt, cz));
    }
}

// ---- shared-memory-tiled variants (3D; selected with LPIC_FDTD_TILED=1, measured slower than the per-cell kernels) -----------------------------------------------------------------------
// One CTA owns FT_X x FT_Y x FT_Z interior cells of one patch.  The three components the curl differentiates are staged in
// shared memory with their one-cell halo (low side for E <- curl B, high side for B <- curl E) in LOGICAL order: every value
// is read from global memory once per tile instead of up to four times through L1, and the wrapped-guard index arithmetic is
// paid per staged value.  The updated component, J and kappa are touched once per cell and stay in global memory.  The
// arithmetic is the per-cell kernels' (same intrinsics, same association), so the result is bit-identical.
constexpr int FT_X = 4, FT_Y = 16, FT_Z = 16, FT_THREADS = 256;  // a 16^3 patch is four such tiles
constexpr int FT_HX = FT_X + 1, FT_HY = FT_Y + 1, FT_HZ = FT_Z + 1;

struct FdtdTile {
    int p, i0, j0, k0;
};
__device__ __forceinline__ FdtdTile fdtd_tile(const Geom &g) {
    const int tz = (g.nz + FT_Z - 1) / FT_Z, ty = (g.ny + FT_Y - 1) / FT_Y, tx = (g.nx + FT_X - 1) / FT_X;
    int b = blockIdx.x;
    FdtdTile t;
    t.k0 = (b % tz) * FT_Z; b /= tz;
    t.j0 = (b % ty) * FT_Y; b /= ty;
    t.i0 = (b % tx) * FT_X;
    t.p = b / tx;
    return t;
}
static unsigned fdtd_tiles(const Geom &g) {
    return (unsigned)((i64)g.npatch * ((g.nx + FT_X - 1) / FT_X) * ((g.ny + FT_Y - 1) / FT_Y) * ((g.nz + FT_Z - 1) / FT_Z));
}

// lo = 1: halo on the low side (staged index 0 is logical origin - 1), lo = 0: halo on the high side
template <int LO>
__device__ __forceinline__ void stage3(const Geom &g, const double *__restrict__ src, const FdtdTile &t, double (*dst)[FT_HY][FT_HZ]) {
    for (int idx = threadIdx.x; idx < FT_HX * FT_HY * FT_HZ; idx += FT_THREADS) {
        const int lz = idx % FT_HZ, ly = (idx / FT_HZ) % FT_HY, lx = idx / (FT_HZ * FT_HY);
        const int i = t.i0 + lx - LO, j = t.j0 + ly - LO, k = t.k0 + lz - LO;
        // (cells past the patch's interior + one halo cell are never read)
        if (i <= g.nx && j <= g.ny && k <= g.nz) dst[lx][ly][lz] = src[sidx(g, i, j, k)];
    }
}

__global__ void __launch_bounds__(FT_THREADS) k_update_efield_tiled(Geom g, double *__restrict__ F, double bfactor, double jfactor,
                                                                    const u8 *__restrict__ is_pml, const double *__restrict__ kappa, int nmax) {
    __shared__ double sbx[FT_HX][FT_HY][FT_HZ], sby[FT_HX][FT_HY][FT_HZ], sbz[FT_HX][FT_HY][FT_HZ];
    const FdtdTile t = fdtd_tile(g);
    const size_t stride = (size_t)g.npatch * g.ncell;
    double *base = F + (size_t)t.p * g.ncell;
    double *ex = base + LPIC_EX * stride, *ey = base + LPIC_EY * stride, *ez = base + LPIC_EZ * stride;
    const double *jx = base + LPIC_JX * stride, *jy = base + LPIC_JY * stride, *jz = base + LPIC_JZ * stride;
    stage3<1>(g, base + LPIC_BX * stride, t, sbx);
    stage3<1>(g, base + LPIC_BY * stride, t, sby);
    stage3<1>(g, base + LPIC_BZ * stride, t, sbz);
    __syncthreads();
    const bool pml = is_pml && is_pml[t.p];
    const double *kp = kappa + (size_t)t.p * 6 * nmax;  // [e|b][axis][nmax], e first
    for (int c = threadIdx.x; c < FT_X * FT_Y * FT_Z; c += FT_THREADS) {
        const int lz = c % FT_Z, ly = (c / FT_Z) % FT_Y, lx = c / (FT_Z * FT_Y);
        const int i = t.i0 + lx, j = t.j0 + ly, k = t.k0 + lz;
        if (i >= g.nx || j >= g.ny || k >= g.nz) continue;
        const int o = sidx(g, i, j, k);
        const double bxc = sbx[lx + 1][ly + 1][lz + 1], byc = sby[lx + 1][ly + 1][lz + 1], bzc = sbz[lx + 1][ly + 1][lz + 1];
        const double bz_ym = sbz[lx + 1][ly][lz + 1], by_zm = sby[lx + 1][ly + 1][lz], bx_zm = sbx[lx + 1][ly + 1][lz];
        const double bz_xm = sbz[lx][ly + 1][lz + 1], by_xm = sby[lx][ly + 1][lz + 1], bx_ym = sbx[lx + 1][ly][lz + 1];
        if (pml) {  // core/boundary/cpml.py:437-457
            const double bfx = __ddiv_rn(bfactor, kp[i]), bfy = __ddiv_rn(bfactor, kp[nmax + j]), bfz = __ddiv_rn(bfactor, kp[2 * nmax + k]);
            const double cx = __dsub_rn(__ddiv_rn(__dmul_rn(bfy, __dsub_rn(bzc, bz_ym)), g.dy), __ddiv_rn(__dmul_rn(bfz, __dsub_rn(byc, by_zm)), g.dz));
            const double cy = __dsub_rn(__ddiv_rn(__dmul_rn(bfz, __dsub_rn(bxc, bx_zm)), g.dz), __ddiv_rn(__dmul_rn(bfx, __dsub_rn(bzc, bz_xm)), g.dx));
            const double cz = __dsub_rn(__ddiv_rn(__dmul_rn(bfx, __dsub_rn(byc, by_xm)), g.dx), __ddiv_rn(__dmul_rn(bfy, __dsub_rn(bxc, bx_ym)), g.dy));
            ex[o] = __dadd_rn(ex[o], __dsub_rn(cx, __dmul_rn(jfactor, jx[o])));
            ey[o] = __dadd_rn(ey[o], __dsub_rn(cy, __dmul_rn(jfactor, jy[o])));
            ez[o] = __dadd_rn(ez[o], __dsub_rn(cz, __dmul_rn(jfactor, jz[o])));
        } else {    // core/maxwell/cpu.py:92-97
            const double cx = __dsub_rn(__ddiv_rn(__dsub_rn(bzc, bz_ym), g.dy), __ddiv_rn(__dsub_rn(byc, by_zm), g.dz));
            const double cy = __dsub_rn(__ddiv_rn(__dsub_rn(bxc, bx_zm), g.dz), __ddiv_rn(__dsub_rn(bzc, bz_xm), g.dx));
            const double cz = __dsub_rn(__ddiv_rn(__dsub_rn(byc, by_xm), g.dx), __ddiv_rn(__dsub_rn(bxc, bx_ym), g.dy));
            ex[o] = __dadd_rn(ex[o], __dsub_rn(__dmul_rn(bfactor, cx), __dmul_rn(jfactor, jx[o])));
            ey[o] = __dadd_rn(ey[o], __dsub_rn(__dmul_rn(bfactor, cy), __dmul_rn(jfactor, jy[o])));
            ez[o] = __dadd_rn(ez[o], __dsub_rn(__dmul_rn(bfactor, cz), __dmul_rn(jfactor, jz[o])));
        }
    }
}

__global__ void __launch_bounds__(FT_THREADS) k_update_bfield_tiled(Geom g, double *__restrict__ F, double dt, const u8 *__restrict__ is_pml,
                                                                    const double *__restrict__ kappa, int nmax) {
    __shared__ double sex[FT_HX][FT_HY][FT_HZ], sey[FT_HX][FT_HY][FT_HZ], sez[FT_HX][FT_HY][FT_HZ];
    const FdtdTile t = fdtd_tile(g);
    const size_t stride = (size_t)g.npatch * g.ncell;
    double *base = F + (size_t)t.p * g.ncell;
    double *bx = base + LPIC_BX * stride, *by = base + LPIC_BY * stride, *bz = base + LPIC_BZ * stride;
    stage3<0>(g, base + LPIC_EX * stride, t, sex);
    stage3<0>(g, base + LPIC_EY * stride, t, sey);
    stage3<0>(g, base + LPIC_EZ * stride, t, sez);
    __syncthreads();
    const bool pml = is_pml && is_pml[t.p];
    const double *kp = kappa + ((size_t)t.p * 6 + 3) * nmax;
    for (int c = threadIdx.x; c < FT_X * FT_Y * FT_Z; c += FT_THREADS) {
        const int lz = c % FT_Z, ly = (c / FT_Z) % FT_Y, lx = c / (FT_Z * FT_Y);
        const int i = t.i0 + lx, j = t.j0 + ly, k = t.k0 + lz;
        if (i >= g.nx || j >= g.ny || k >= g.nz) continue;
        const int o = sidx(g, i, j, k);
        const double exc = sex[lx][ly][lz], eyc = sey[lx][ly][lz], ezc = sez[lx][ly][lz];
        const double ez_yp = sez[lx][ly + 1][lz], ey_zp = sey[lx][ly][lz + 1], ex_zp = sex[lx][ly][lz + 1];
        const double ez_xp = sez[lx + 1][ly][lz], ey_xp = sey[lx + 1][ly][lz], ex_yp = sex[lx][ly + 1][lz];
        if (pml) {  // core/boundary/cpml.py:459-477
            const double efx = __ddiv_rn(dt, kp[i]), efy = __ddiv_rn(dt, kp[nmax + j]), efz = __ddiv_rn(dt, kp[2 * nmax + k]);
            const double cx = __dsub_rn(__ddiv_rn(__dmul_rn(efy, __dsub_rn(ez_yp, ezc)), g.dy), __ddiv_rn(__dmul_rn(efz, __dsub_rn(ey_zp, eyc)), g.dz));
            const double cy = __dsub_rn(__ddiv_rn(__dmul_rn(efz, __dsub_rn(ex_zp, exc)), g.dz), __ddiv_rn(__dmul_rn(efx, __dsub_rn(ez_xp, ezc)), g.dx));
            const double cz = __dsub_rn(__ddiv_rn(__dmul_rn(efx, __dsub_rn(ey_xp, eyc)), g.dx), __ddiv_rn(__dmul_rn(efy, __dsub_rn(ex_yp, exc)), g.dy));
            bx[o] = __dsub_rn(bx[o], cx);
            by[o] = __dsub_rn(by[o], cy);
            bz[o] = __dsub_rn(bz[o], cz);
        } else {    // core/maxwell/cpu.py:131-136
            const double cx = __dsub_rn(__ddiv_rn(__dsub_rn(ez_yp, ezc), g.dy), __ddiv_rn(__dsub_rn(ey_zp, eyc), g.dz));
            const double cy = __dsub_rn(__ddiv_rn(__dsub_rn(ex_zp, exc), g.dz), __ddiv_rn(__dsub_rn(ez_xp, ezc), g.dx));
            const double cz = __dsub_rn(__ddiv_rn(__dsub_rn(ey_xp, eyc), g.dx), __ddiv_rn(__dsub_rn(ex_yp, exc), g.dy));
            bx[o] = __dsub_rn(bx[o], __dmul_rn(dt, cx));
            by[o] = __dsub_rn(by[o], __dmul_rn(dt, cy));
            bz[o] = __dsub_rn(bz[o], __dmul_rn(dt, cz));
        }
    }
}

// psi update of the PML faces in one slot and the correction of the two components each face drives
// (core/boundary/cpml.py:527-730).  grid.y = instance inside the slot; one thread per interior cell of its patch.
//   E: x: (ey,-,bz) (ez,+,by)   y: (ex,+,bz) (ez,-,bx)   z: (ex,-,by) (ey,+,bx)      backward difference of B, fac = dt c^2
//   B: x: (by,+,ez) (bz,-,ey)   y: (bx,-,ez) (bz,+,ex)   z: (bx,+,ey) (by,-,ex)      forward difference of E,  fac = dt
__global__ void __launch_bounds__(256) k_pml_psi(Geom g, double *__restrict__ F, const int *__restrict__ inst, const i64 *__restrict__ order,
                                                 i64 first, const double *__restrict__ coef, double *__restrict__ psi, int nmax,
                                                 i64 ncint, int is_b, double fac) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= g.nx * g.ny * g.nz) return;
    const i64 e = first + blockIdx.y;
    const int *in = inst + e * 8;
    const int p = in[0], axis = in[1], lo = in[2 + 2 * is_b], hi = in[3 + 2 * is_b];
    const int k = t % g.nz, j = (t / g.nz) % g.ny, i = t / (g.nz * g.ny);
    const int ipos = axis == 0 ? i : (axis == 1 ? j : k);
    if (ipos < lo || ipos >= hi) return;
    // component table: index into the field arena (LPIC_EX.. LPIC_BZ)
    int c1, c2, g1, g2;
    double s1, s2;
    if (!is_b) {
        if (axis == 0) { c1 = LPIC_EY; s1 = -1; g1 = LPIC_BZ; c2 = LPIC_EZ; s2 = 1; g2 = LPIC_BY; }
        else if (axis == 1) { c1 = LPIC_EX; s1 = 1; g1 = LPIC_BZ; c2 = LPIC_EZ; s2 = -1; g2 = LPIC_BX; }
        else { c1 = LPIC_EX; s1 = -1; g1 = LPIC_BY; c2 = LPIC_EY; s2 = 1; g2 = LPIC_BX; }
    } else {
        if (axis == 0) { c1 = LPIC_BY; s1 = 1; g1 = LPIC_EZ; c2 = LPIC_BZ; s2 = -1; g2 = LPIC_EY; }
        else if (axis == 1) { c1 = LPIC_BX; s1 = -1; g1 = LPIC_EZ; c2 = LPIC_BZ; s2 = 1; g2 = LPIC_EX; }
        else { c1 = LPIC_BX; s1 = 1; g1 = LPIC_EY; c2 = LPIC_BY; s2 = -1; g2 = LPIC_EX; }
    }
    const size_t stride = (size_t)g.npatch * g.ncell;
    double *base = F + (size_t)p * g.ncell;
    const int o = sidx(g, i, j, k);
    const int d = is_b ? 1 : -1;
    const int nbr = sidx(g, i + (axis == 0 ? d : 0), j + (axis == 1 ? d : 0), k + (axis == 2 ? d : 0));
    const double bco = coef[((size_t)e * 4 + 2 * is_b) * nmax + ipos], cco = coef[((size_t)e * 4 + 2 * is_b + 1) * nmax + ipos];
    const double a1 = base[g1 * stride + o], n1 = base[g1 * stride + nbr], a2 = base[g2 * stride + o], n2 = base[g2 * stride + nbr];
    const double d1 = is_b ? __dsub_rn(n1, a1) : __dsub_rn(a1, n1), d2 = is_b ? __dsub_rn(n2, a2) : __dsub_rn(a2, n2);
    double *ps = psi + ((size_t)order[e] * 4 + 2 * is_b) * ncint + t;
    const double p1 = __dadd_rn(__dmul_rn(bco, ps[0]), __dmul_rn(cco, d1));
    const double p2 = __dadd_rn(__dmul_rn(bco, ps[ncint]), __dmul_rn(cco, d2));
    ps[0] = p1;
    ps[ncint] = p2;
    base[c1 * stride + o] = __dadd_rn(base[c1 * stride + o], __dmul_rn(s1, __dmul_rn(fac, p1)));
    base[c2 * stride + o] = __dadd_rn(base[c2 * stride + o], __dmul_rn(s2, __dmul_rn(fac, p2)));
}

// Laser antenna (callback/laser.py:17-77): one thread per (listed patch, y, z) of the antenna plane.  Every product and sum
// is an explicit intrinsic in the reference's left-to-right order, so the result is bit-identical to the numba kernel.
__global__ void __launch_bounds__(128) k_laser(Geom g, double *__restrict__ F, const int *__restrict__ patches, const int *__restrict__ ranges,
                                               const double *__restrict__ ey_src, const double *__restrict__ ez_src, int laserpos,
                                               double dt) {
    const int e = blockIdx.y;
    const int p = patches[e];
    const int *rg = ranges + 4 * e;
    const int plane = g.NY * g.NZ;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= plane) return;
    const int sk = t % g.NZ, sj = t / g.NZ;
    const int iy = sj < g.ny + g.ng ? sj : sj - g.NY, iz = sk < g.nz + g.ngz ? sk : sk - g.NZ;  // logical indices
    if (iy < rg[0] || iy >= rg[1]) return;
    const size_t stride = (size_t)g.npatch * g.ncell;
    double *base = F + (size_t)p * g.ncell;
    const double *ey = base + LPIC_EY * stride, *ez = base + LPIC_EZ * stride, *jy = base + LPIC_JY * stride, *jz = base + LPIC_JZ * stride;
    double *bx = base + LPIC_BX * stride, *by = base + LPIC_BY * stride, *bz = base + LPIC_BZ * stride;
    // bx[laserpos-1, iy, :] = bx[0, iy, :]  -- the whole padded z row
    bx[sidx(g, laserpos - 1, iy, 0) + sk] = bx[sidx(g, 0, iy, 0) + sk];
    if (g.dim == 3 && (iz < rg[2] || iz >= rg[3])) return;
    if (g.dim == 2 && sk != 0) return;
    const double c = LPIC_C_LIGHT;
    const int o0 = sidx(g, 0, iy, iz), om = sidx(g, -1, iy, iz), ol = sidx(g, laserpos, iy, iz), ot = sidx(g, laserpos - 1, iy, iz);
    const double cdtdx = __ddiv_rn(__dmul_rn(c, dt), g.dx);
    const double inv = __ddiv_rn(1.0, __dmul_rn(__dadd_rn(cdtdx, 1.0), c));
    const double dteps = __ddiv_rn(dt, LPIC_EPS0), dtc2 = __dmul_rn(dt, __dmul_rn(c, c)), back = __dmul_rn(__dsub_rn(cdtdx, 1.0), c);
    const double half_c = __dmul_rn(c, 0.5);
    const size_t s = (size_t)e * plane + t;
    double vz = __dadd_rn(__dmul_rn(4.0, ey_src[s]), __dmul_rn(2.0, __dadd_rn(ey[o0], __dmul_rn(half_c, __dadd_rn(bz[o0], bz[om])))));
    vz = __dsub_rn(vz, __dmul_rn(2.0, ey[ol]));
    if (g.dim == 3) vz = __dsub_rn(vz, __ddiv_rn(__dmul_rn(dtc2, __dsub_rn(bx[ol], bx[sidx(g, laserpos, iy, iz - 1)])), g.dz));
    vz = __dadd_rn(vz, __dmul_rn(dteps, jy[ol]));
    vz = __dadd_rn(vz, __dmul_rn(back, bz[ol]));
    double vy = __dsub_rn(__dmul_rn(-4.0, ez_src[s]), __dmul_rn(2.0, __dsub_rn(ez[o0], __dmul_rn(half_c, __dadd_rn(by[o0], by[om])))));
    vy = __dadd_rn(vy, __dmul_rn(2.0, ez[ol]));
    vy = __dsub_rn(vy, __ddiv_rn(__dmul_rn(dtc2, __dsub_rn(bx[ol], bx[sidx(g, laserpos, iy - 1, iz)])), g.dy));
    vy = __dsub_rn(vy, __dmul_rn(dteps, jz[ol]));
    vy = __dadd_rn(vy, __dmul_rn(back, by[ol]));
    bz[ot] = __dmul_rn(inv, vz);
    by[ot] = __dmul_rn(inv, vy);
}

// storage index -> logical index along one axis
__device__ __forceinline__ int logical(int s, int n, int ng) { return s < n + ng ? s : s - (n + 2 * ng); }

// Guard copy: one thread per (attribute, patch, padded cell); guard cells pull from the neighbour's interior strip.
// dst[-ng,0) <- src[n-ng,n), dst[n,n+ng) <- src[0,ng) per axis (sync_fields2d.c:191-196).  Every guard cell has
// exactly one source, so the order of boundaries in the reference does not matter for a copy.
struct AttrList {
    int n;
    int a[LPIC_NFIELD];
};
__global__ void __launch_bounds__(256) k_sync_guard(Geom g, double *__restrict__ F, const i64 *__restrict__ nbr,
                                                    AttrList attrs) {
    // direction -> boundary table in shared memory: lanes of a warp look up different entries, which would serialise
    // in the constant cache
    __shared__ signed char lut[27];
    if (threadIdx.x < 27) lut[threadIdx.x] = g.dim == 3 ? kLut3[threadIdx.x] : (threadIdx.x < 9 ? kLut2[threadIdx.x] : -1);
    __syncthreads();
    const i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (i64)g.npatch * g.ncell) return;
    const i64 t0 = t - threadIdx.x;  // block-uniform 64-bit division, 32-bit arithmetic per lane
    int p = (int)(t0 / g.ncell);
    int r = (int)(t0 - (i64)p * g.ncell) + (int)threadIdx.x;
    if (r >= g.ncell) { p += r / g.ncell; r %= g.ncell; }
    const int sk = r % g.NZ;
    r /= g.NZ;
    const int sj = r % g.NY, si = r / g.NY;
    const int li = logical(si, g.nx, g.ng), lj = logical(sj, g.ny, g.ng), lk = logical(sk, g.nz, g.ngz);
    const int sx = li < 0 ? -1 : (li >= g.nx ? 1 : 0), sy = lj < 0 ? -1 : (lj >= g.ny ? 1 : 0),
              sz = lk < 0 ? -1 : (lk >= g.nz ? 1 : 0);
    if (sx == 0 && sy == 0 && sz == 0) return;
    const int b = lut[(sx + 1) + 3 * (sy + 1) + (g.dim == 3 ? 9 * (sz + 1) : 0)];
    const i64 q = nbr[(i64)p * g.nb + b];
    if (q < 0) return;
    const int qi = li - sx * g.nx, qj = lj - sy * g.ny, qk = lk - sz * g.nz;  // interior of the neighbour
    const int src = qk + g.NZ * (qj + g.NY * qi);
    const int dst = sk + g.NZ * (sj + g.NY * si);
    // every requested component in the same thread: the index arithmetic is paid once and the loads are independent
    const size_t stride = (size_t)g.npatch * g.ncell, from = (size_t)q * g.ncell + src, to = (size_t)p * g.ncell + dst;
    double v[LPIC_NFIELD];
#pragma unroll
    for (int ai = 0; ai < LPIC_NFIELD; ai++)
        if (ai < attrs.n) v[ai] = F[attrs.a[ai] * stride + from];
#pragma unroll
    for (int ai = 0; ai < LPIC_NFIELD; ai++)
        if (ai < attrs.n) F[attrs.a[ai] * stride + to] = v[ai];
}

// bit masks of the boundaries whose direction component along an axis is -1 / 0 / +1 (generated from kDir3 / kDir2)
__constant__ const unsigned kAxisMask3[3][3] = {{0x3c03c1u, 0x3c03cu, 0x3c03c02u}, {0xccc444u, 0x3333u, 0x3330888u}, {0x1555110u, 0xccfu, 0x2aaa220u}};
__constant__ const unsigned kAxisMask2[3][3] = {{0x51u, 0xcu, 0xa2u}, {0x34u, 0x3u, 0xc8u}, {0x0u, 0xffu, 0x0u}};

// Current reduce: one thread per (attribute, patch, interior cell).  The thread walks the boundaries in the
// reference's enum order (faces, edges, vertices), adds the neighbour's guard value and zeroes it, exactly as
// sync_currents_3d does cell by cell (sync_fields3d.c:117-128), which fixes the summation order at edge and
// corner cells.  The consumed guard cells are zeroed by a second streaming kernel (k_zero_consumed_guards): mixing the
// scattered zero stores with the dependent neighbour loads in one kernel made every load wait ~14k cycles (ncu:
// long_scoreboard 269 cycles per issue at 2 % DRAM utilisation).
__global__ void __launch_bounds__(256) k_sync_currents(Geom g, double *__restrict__ F, const i64 *__restrict__ nbr) {
    __shared__ signed char sdir[26][3];  // per-lane lookups by boundary id: shared memory, not the constant cache
    if (threadIdx.x < 26 * 3) {
        const int b = threadIdx.x / 3, ax = threadIdx.x - 3 * b;
        sdir[b][ax] = b < g.nb ? (g.dim == 3 ? kDir3[b][ax] : kDir2[b][ax]) : 0;
    }
    __syncthreads();
    CellIdx c;
    if (!interior_cell(g, (i64)blockIdx.x * blockDim.x + threadIdx.x, c)) return;
    const bool lox = c.i < g.ng, hix = c.i >= g.nx - g.ng;
    const bool loy = c.j < g.ng, hiy = c.j >= g.ny - g.ng;
    const bool loz = g.dim == 3 && c.k < g.ng, hiz = g.dim == 3 && c.k >= g.nz - g.ng;
    if (!(lox || hix || loy || hiy || loz || hiz)) return;
    double *base = F + (size_t)(LPIC_JX + blockIdx.y) * g.npatch * g.ncell;
    const int o = c.k + g.NZ * (c.j + g.NY * c.i);
    double acc = base[(size_t)c.p * g.ncell + o];
    bool touched = false;
    // boundaries whose strip contains this cell, as a bit mask, visited in ascending (= enum) order
    const unsigned(*am)[3] = g.dim == 3 ? kAxisMask3 : kAxisMask2;
    unsigned todo = (am[0][1] | (lox ? am[0][0] : 0u) | (hix ? am[0][2] : 0u)) &
                    (am[1][1] | (loy ? am[1][0] : 0u) | (hiy ? am[1][2] : 0u)) &
                    (am[2][1] | (loz ? am[2][0] : 0u) | (hiz ? am[2][2] : 0u));
    while (todo) {
        const int b = __ffs(todo) - 1;
        todo &= todo - 1;
        const i64 q = nbr[(i64)c.p * g.nb + b];
        if (q < 0) continue;
        const int sx = sdir[b][0], sy = sdir[b][1], sz = sdir[b][2];
        // dst[0,ng) += src[n,n+ng) for a MIN side, dst[n-ng,n) += src[-ng,0) for a MAX side
        const int qi = c.i - sx * g.nx, qj = c.j - sy * g.ny, qk = c.k - sz * g.nz;
        const size_t s = (size_t)q * g.ncell + sidx(g, qi, qj, qk);
        acc = __dadd_rn(acc, base[s]);
        touched = true;
    }
    if (touched) base[(size_t)c.p * g.ncell + o] = acc;
}

// zero every guard cell whose neighbour in that direction exists (that neighbour has just reduced it)
__global__ void __launch_bounds__(256) k_zero_consumed_guards(Geom g, double *__restrict__ F, const i64 *__restrict__ nbr) {
    __shared__ signed char lut[27];
    if (threadIdx.x < 27) lut[threadIdx.x] = g.dim == 3 ? kLut3[threadIdx.x] : (threadIdx.x < 9 ? kLut2[threadIdx.x] : -1);
    __syncthreads();
    const i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (i64)g.npatch * g.ncell) return;
    const i64 t0 = t - threadIdx.x;
    int p = (int)(t0 / g.ncell);
    int r = (int)(t0 - (i64)p * g.ncell) + (int)threadIdx.x;
    if (r >= g.ncell) { p += r / g.ncell; r %= g.ncell; }
    const int cell = r;
    const int sk = r % g.NZ;
    r /= g.NZ;
    const int sj = r % g.NY, si = r / g.NY;
    const int li = logical(si, g.nx, g.ng), lj = logical(sj, g.ny, g.ng), lk = logical(sk, g.nz, g.ngz);
    const int sx = li < 0 ? -1 : (li >= g.nx ? 1 : 0), sy = lj < 0 ? -1 : (lj >= g.ny ? 1 : 0),
              sz = lk < 0 ? -1 : (lk >= g.nz ? 1 : 0);
    if (sx == 0 && sy == 0 && sz == 0) return;
    const int b = lut[(sx + 1) + 3 * (sy + 1) + (g.dim == 3 ? 9 * (sz + 1) : 0)];
    if (nbr[(i64)p * g.nb + b] < 0) return;
    const size_t stride = (size_t)g.npatch * g.ncell;
    double *base = F + (size_t)LPIC_JX * stride + (size_t)p * g.ncell + cell;
    base[0] = 0.0; base[stride] = 0.0; base[2 * stride] = 0.0; base[3 * stride] = 0.0;
}

__global__ void __launch_bounds__(256) k_field_energy(Geom g, const double *__restrict__ F, double *__restrict__ out) {
    CellIdx c;
    double e2 = 0.0, b2 = 0.0;
    for (i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x; interior_cell(g, t, c); t += (i64)gridDim.x * blockDim.x) {
        const size_t stride = (size_t)g.npatch * g.ncell;
        const double *base = F + (size_t)c.p * g.ncell + sidx(g, c.i, c.j, c.k);
        double a;
        a = base[LPIC_EX * stride]; e2 += a * a;
        a = base[LPIC_EY * stride]; e2 += a * a;
        a = base[LPIC_EZ * stride]; e2 += a * a;
        a = base[LPIC_BX * stride]; b2 += a * a;
        a = base[LPIC_BY * stride]; b2 += a * a;
        a = base[LPIC_BZ * stride]; b2 += a * a;
    }
    for (int o = 16; o; o >>= 1) {
        e2 += __shfl_xor_sync(0xffffffffu, e2, o);
        b2 += __shfl_xor_sync(0xffffffffu, b2, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(out, e2);
        atomicAdd(out + 1, b2);
    }
}

}  // namespace

// psi coefficients for this dt (bcoeff = exp(-(sigma/kappa + a) dt), ccoeff = (bcoeff - 1) sigma / kappa / (sigma + kappa a) / d,
// core/boundary/cpml.py:535-536), then one psi launch per slot so that the faces of a patch are applied in list order
static int pml_advance(lpic_ctx *c, int is_b, double dt) {
    PmlState *pm = c->pml;
    const Geom &g = c->g;
    if (pm->ninst == 0) return 0;
    if (pm->coef_dt[is_b] != dt) {
        std::vector<double> co((size_t)pm->ninst * 2 * pm->nmax, 0.0);
        const double dd[3] = {g.dx, g.dy, g.dz};
        for (i64 e = 0; e < pm->ninst; e++) {
            const double *kap = pm->h_prof + ((size_t)e * 6 + 3 * is_b) * pm->nmax, *sig = kap + pm->nmax, *a = sig + pm->nmax;
            for (i64 i = 0; i < pm->nmax; i++) {
                const double bco = exp(-(sig[i] / kap[i] + a[i]) * dt);
                const double den = sig[i] + kap[i] * a[i];
                co[((size_t)e * 2) * pm->nmax + i] = bco;
                co[((size_t)e * 2 + 1) * pm->nmax + i] = den != 0.0 ? (bco - 1) * sig[i] / kap[i] / den / dd[pm->h_axis[e]] : 0.0;
            }
        }
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        for (i64 e = 0; e < pm->ninst; e++)
            CUDA_TRY(cudaMemcpy(pm->d_coef + ((size_t)e * 4 + 2 * is_b) * pm->nmax, co.data() + (size_t)e * 2 * pm->nmax,
                                sizeof(double) * 2 * pm->nmax, cudaMemcpyHostToDevice));
        pm->coef_dt[is_b] = dt;
    }
    const double fac = is_b ? dt : dt * (LPIC_C_LIGHT * LPIC_C_LIGHT);
    for (int sl = 0; sl < pm->nslot; sl++) {
        const i64 cnt = pm->slot_first[sl + 1] - pm->slot_first[sl];
        if (cnt == 0) continue;
        dim3 grid(div_up((i64)g.nx * g.ny * g.nz, 256), (unsigned)cnt);
        k_pml_psi<<<grid, 256, 0, c->stream>>>(g, c->fields, pm->d_inst, (const i64 *)pm->d_order, pm->slot_first[sl], pm->d_coef,
                                               pm->d_psi, (int)pm->nmax, pm->ncint, is_b, fac);
        LAUNCHED(1);
    }
    KERNEL_CHECK();
    return 0;
}

extern "C" int lpic_update_efield(lpic_ctx *c, double dt) {
    DeviceGuard dg(c);
    const Geom &g = c->g;
    const i64 n = (i64)g.npatch * g.nx * g.ny * g.nz;
    // bfactor = dt*c**2, jfactor = dt/epsilon_0 (core/maxwell/cpu.py:90-91)
    const double bfactor = dt * (LPIC_C_LIGHT * LPIC_C_LIGHT), jfactor = dt / LPIC_EPS0;
    const PmlState *pm = c->pml;
    const u8 *isp = pm ? pm->d_is_pml : nullptr;
    const double *kap = pm ? pm->d_kappa : nullptr;
    const int nmax = pm ? (int)pm->nmax : 0;
    // A/B (256^3 cells, one B200, profiles/r2_fdtd_tiled_ab.txt): the one-thread-per-cell kernel 0.707 ms, the shared-memory tile
    // 0.818 ms -- every B value is reused at most four times and L1 already serves those re-reads, so the staging pass and its
    // barrier cost more than they save.  The per-cell kernel stays the default; LPIC_FDTD_TILED=1 selects the tiled one.
    static const bool per_cell = getenv("LPIC_FDTD_TILED") == nullptr;
    if (g.dim == 3 && !per_cell)
        k_update_efield_tiled<<<fdtd_tiles(g), FT_THREADS, 0, c->stream>>>(g, c->fields, bfactor, jfactor, isp, kap, nmax);
    else if (g.dim == 3)
        k_update_efield<3><<<div_up(n, 256), 256, 0, c->stream>>>(g, c->fields, bfactor, jfactor, isp, kap, nmax);
    else
        k_update_efield<2><<<div_up(n, 256), 256, 0, c->stream>>>(g, c->fields, bfactor, jfactor, isp, kap, nmax);
    LAUNCHED(1);
    KERNEL_CHECK();
    return pm ? pml_advance(c, 0, dt) : 0;
}

extern "C" int lpic_update_bfield(lpic_ctx *c, double dt) {
    DeviceGuard dg(c);
    const Geom &g = c->g;
    const i64 n = (i64)g.npatch * g.nx * g.ny * g.nz;
    const PmlState *pm = c->pml;
    const u8 *isp = pm ? pm->d_is_pml : nullptr;
    const double *kap = pm ? pm->d_kappa : nullptr;
    const int nmax = pm ? (int)pm->nmax : 0;
    static const bool per_cell = getenv("LPIC_FDTD_TILED") == nullptr;  // see lpic_update_efield: 0.526 ms per-cell, 0.602 ms tiled
    if (g.dim == 3 && !per_cell)
        k_update_bfield_tiled<<<fdtd_tiles(g), FT_THREADS, 0, c->stream>>>(g, c->fields, dt, isp, kap, nmax);
    else if (g.dim == 3)
        k_update_bfield<3><<<div_up(n, 256), 256, 0, c->stream>>>(g, c->fields, dt, isp, kap, nmax);
    else
        k_update_bfield<2><<<div_up(n, 256), 256, 0, c->stream>>>(g, c->fields, dt, isp, kap, nmax);
    LAUNCHED(1);
    KERNEL_CHECK();
    return pm ? pml_advance(c, 1, dt) : 0;
}

extern "C" int lpic_sync_guard_fields(lpic_ctx *c, uint32_t attr_mask) {
    DeviceGuard dg(c);
    const Geom &g = c->g;
    AttrList attrs;
    int na = 0;
    for (int a = 0; a < LPIC_NFIELD; a++)
        if (attr_mask & (1u << a)) attrs.a[na++] = a;
    if (!na) return 0;
    attrs.n = na;
    const unsigned grid = (unsigned)div_up((i64)g.npatch * g.ncell, 256);
    k_sync_guard<<<grid, 256, 0, c->stream>>>(g, c->fields, c->d_nbr, attrs);
    LAUNCHED(1);
    KERNEL_CHECK();
    return 0;
}

extern "C" int lpic_sync_currents(lpic_ctx *c) {
    DeviceGuard dg(c);
    const Geom &g = c->g;
    dim3 grid(div_up((i64)g.npatch * g.nx * g.ny * g.nz, 256), 4);
    k_sync_currents<<<grid, 256, 0, c->stream>>>(g, c->fields, c->d_nbr);
    k_zero_consumed_guards<<<div_up((i64)g.npatch * g.ncell, 256), 256, 0, c->stream>>>(g, c->fields, c->d_nbr);
    LAUNCHED(2);
    KERNEL_CHECK();
    return 0;
}

extern "C" int lpic_reset_currents(lpic_ctx *c) {
    DeviceGuard dg(c);
    const Geom &g = c->g;
    CUDA_TRY(cudaMemsetAsync(field_ptr(c, LPIC_JX), 0, sizeof(double) * 4 * (size_t)g.npatch * g.ncell, c->stream));
    return 0;
}

extern "C" int lpic_field_energy_sums(lpic_ctx *c, double *out2) {
    DeviceGuard dg(c);
    const Geom &g = c->g;
    CUDA_TRY(cudaMemsetAsync(c->d_tmpf, 0, 2 * sizeof(double), c->stream));
    k_field_energy<<<148 * 4, 256, 0, c->stream>>>(g, c->fields, c->d_tmpf);
    LAUNCHED(1);
    KERNEL_CHECK();
    CUDA_TRY(cudaMemcpyAsync(out2, c->d_tmpf, 2 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return 0;
}

void lpic_free_pml(lpic_ctx *c) {
    PmlState *pm = c->pml;
    if (!pm) return;
    cudaFree(pm->d_is_pml); cudaFree(pm->d_kappa); cudaFree(pm->d_inst); cudaFree(pm->d_coef); cudaFree(pm->d_psi); cudaFree(pm->d_order);
    delete[] pm->h_prof; delete[] pm->h_order; delete[] pm->h_axis;
    delete pm;
    c->pml = nullptr;
}

extern "C" int lpic_pml_configure(lpic_ctx *c, int64_t ninst, const int64_t *inst_patch, const int64_t *inst_axis,
                                  const int64_t *inst_slot, const int64_t *ranges, const double *profiles, int64_t nmax) {
    DeviceGuard dg(c);
    const Geom &g = c->g;
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    lpic_free_pml(c);
    if (ninst <= 0) return 0;
    REQUIRE(nmax >= g.nx && nmax >= g.ny && nmax >= g.nz, "nmax must cover the longest patch axis");
    PmlState *pm = new PmlState();
    c->pml = pm;
    pm->ninst = ninst; pm->nmax = nmax; pm->ncint = (i64)g.nx * g.ny * g.nz;
    // group the instances by slot (the order in which a patch applies its faces)
    std::vector<i64> order;
    pm->nslot = 0;
    for (i64 e = 0; e < ninst; e++) {
        REQUIRE(inst_slot[e] >= 0 && inst_slot[e] < 4 && inst_axis[e] >= 0 && inst_axis[e] < g.dim && inst_patch[e] >= 0 &&
                    inst_patch[e] < g.npatch, "bad PML instance %lld", (long long)e);
        pm->nslot = std::max(pm->nslot, (int)inst_slot[e] + 1);
    }
    for (int sl = 0; sl < pm->nslot; sl++) {
        pm->slot_first[sl] = (i64)order.size();
        for (i64 e = 0; e < ninst; e++)
            if (inst_slot[e] == sl) order.push_back(e);
    }
    for (int sl = pm->nslot; sl <= 4; sl++) pm->slot_first[sl] = (i64)order.size();
    pm->h_order = new i64[ninst]; pm->h_axis = new int[ninst];
    pm->h_prof = new double[(size_t)ninst * 6 * nmax];
    std::vector<int> inst((size_t)ninst * 8, 0);
    std::vector<u8> isp(g.npatch, 0);
    std::vector<double> kap((size_t)g.npatch * 6 * nmax, 1.0);
    for (i64 s = 0; s < ninst; s++) {
        const i64 e = order[s];
        pm->h_order[s] = e;
        pm->h_axis[s] = (int)inst_axis[e];
        memcpy(pm->h_prof + (size_t)s * 6 * nmax, profiles + (size_t)e * 6 * nmax, sizeof(double) * 6 * nmax);
        int *in = inst.data() + s * 8;
        in[0] = (int)inst_patch[e]; in[1] = (int)inst_axis[e];
        for (int r = 0; r < 4; r++) in[2 + r] = (int)ranges[e * 4 + r];
        isp[inst_patch[e]] = 1;
        // this face's kappa profile replaces the default 1.0 along its axis (solver.py:88-106)
        for (int w = 0; w < 2; w++)
            memcpy(kap.data() + (((size_t)inst_patch[e] * 2 + w) * 3 + inst_axis[e]) * nmax, profiles + ((size_t)e * 6 + 3 * w) * nmax,
                   sizeof(double) * nmax);
    }
    CUDA_TRY(cudaMalloc(&pm->d_is_pml, g.npatch));
    CUDA_TRY(cudaMemcpy(pm->d_is_pml, isp.data(), g.npatch, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMalloc(&pm->d_kappa, sizeof(double) * kap.size()));
    CUDA_TRY(cudaMemcpy(pm->d_kappa, kap.data(), sizeof(double) * kap.size(), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMalloc(&pm->d_inst, sizeof(int) * inst.size()));
    CUDA_TRY(cudaMemcpy(pm->d_inst, inst.data(), sizeof(int) * inst.size(), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMalloc(&pm->d_order, sizeof(i64) * ninst));
    CUDA_TRY(cudaMemcpy(pm->d_order, pm->h_order, sizeof(i64) * ninst, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMalloc(&pm->d_coef, sizeof(double) * (size_t)ninst * 4 * nmax));
    CUDA_TRY(cudaMemset(pm->d_coef, 0, sizeof(double) * (size_t)ninst * 4 * nmax));
    CUDA_TRY(cudaMalloc(&pm->d_psi, sizeof(double) * (size_t)ninst * 4 * pm->ncint));
    CUDA_TRY(cudaMemset(pm->d_psi, 0, sizeof(double) * (size_t)ninst * 4 * pm->ncint));
    return 0;
}

// psi arrays of every CPML face that belongs to one of the listed patches -> 0 (the patch was recycled by the moving window)
int lpic_pml_zero_psi_of(lpic_ctx *c, i64 n, const int64_t *patches) {
    PmlState *pm = c->pml;
    if (!pm || pm->ninst == 0) return 0;
    std::vector<int> inst((size_t)pm->ninst * 8);
    CUDA_TRY(cudaMemcpyAsync(inst.data(), pm->d_inst, sizeof(int) * inst.size(), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    for (i64 s = 0; s < pm->ninst; s++)
        for (i64 i = 0; i < n; i++)
            if (inst[s * 8] == (int)patches[i])  // slot-grouped position s holds the caller's instance h_order[s]
                CUDA_TRY(cudaMemsetAsync(pm->d_psi + (size_t)pm->h_order[s] * 4 * pm->ncint, 0, sizeof(double) * 4 * pm->ncint, c->stream));
    return 0;
}

extern "C" int64_t lpic_pml_psi_words(const lpic_ctx *c) {
    DeviceGuard dg(c); return c->pml ? c->pml->ninst * 4 * c->pml->ncint : 0; }
extern "C" int lpic_pml_upload_psi(lpic_ctx *c, const double *host) {
    DeviceGuard dg(c);
    if (!c->pml) return 0;
    CUDA_TRY(cudaMemcpyAsync(c->pml->d_psi, host, sizeof(double) * (size_t)lpic_pml_psi_words(c), cudaMemcpyHostToDevice, c->stream));
    return 0;
}
extern "C" int lpic_pml_download_psi(lpic_ctx *c, double *host) {
    DeviceGuard dg(c);
    if (!c->pml) return 0;
    CUDA_TRY(cudaMemcpyAsync(host, c->pml->d_psi, sizeof(double) * (size_t)lpic_pml_psi_words(c), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int lpic_laser_bfields(lpic_ctx *c, int64_t laserpos, int64_t n, const int64_t *patches, const int64_t *ranges,
                                  const double *ey_src, const double *ez_src, double dt) {
    DeviceGuard dg(c);
    const Geom &g = c->g;
    if (n <= 0) return 0;
    REQUIRE(laserpos >= 1 && laserpos < g.nx + g.ng, "laserpos %lld outside the padded patch", (long long)laserpos);
    const size_t plane = (size_t)g.NY * g.NZ;
    std::vector<int> hp(n), hr(4 * n);
    for (i64 e = 0; e < n; e++) {
        REQUIRE(patches[e] >= 0 && patches[e] < g.npatch, "bad patch in the laser list");
        hp[e] = (int)patches[e];
        for (int r = 0; r < 4; r++) hr[4 * e + r] = (int)ranges[4 * e + r];
    }
    // persistent staging, grown on demand; the source planes come from pageable host memory, so the copies below are
    // synchronous with respect to the host and the buffers may be reused by the next call without a stream synchronisation
    if ((size_t)n > c->laser_cap_n || 2 * (size_t)n * plane > c->laser_cap_words) {
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        cudaFree(c->d_laser_i); cudaFree(c->d_laser_s);
        c->d_laser_i = nullptr; c->d_laser_s = nullptr; c->laser_cap_n = 0; c->laser_cap_words = 0;
        CUDA_TRY(cudaMalloc(&c->d_laser_i, sizeof(int) * 5 * n));
        CUDA_TRY(cudaMalloc(&c->d_laser_s, sizeof(double) * 2 * n * plane));
        c->laser_cap_n = (size_t)n; c->laser_cap_words = 2 * (size_t)n * plane;
    }
    int *d_i = c->d_laser_i;
    double *d_s = c->d_laser_s;
    CUDA_TRY(cudaMemcpyAsync(d_i, hp.data(), sizeof(int) * n, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(d_i + n, hr.data(), sizeof(int) * 4 * n, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(d_s, ey_src, sizeof(double) * n * plane, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(d_s + n * plane, ez_src, sizeof(double) * n * plane, cudaMemcpyHostToDevice, c->stream));
    dim3 grid(div_up((i64)plane, 128), (unsigned)n);
    k_laser<<<grid, 128, 0, c->stream>>>(g, c->fields, d_i, d_i + n, d_s, d_s + n * plane, (int)laserpos, dt);
    LAUNCHED(1);
    KERNEL_CHECK();
    return 0;
}
