// Field gather (quadratic spline on the staggered Yee grid) + relativistic Boris push + Esirkepov deposit.
//
// Reference behaviour restated (not copied): core/pusher/unified/unified_pusher_3d.c:15-217,281-431 and
// unified_pusher_2d.c:64-365 (half push, gather, Boris, half push), core/current/current_deposit.h:7-35,
// 150-268 (2D), 275-440 (3D) (charge-conserving deposit), core/interpolation/cpu3d.c:51-97,
// core/pusher/boris.py:6-38, core/pusher/cpu.py:58-90.
//
// Kernel structure (round 1): one thread per particle slot, blocks never straddle a patch so the patch's
// grid base pointers are block-uniform; SoA attribute streams are read/written coalesced; E/B are read
// through L1 (a patch tile + guards is 85 KB per component at 16^3); J/rho are accumulated with native
// fp64 global reductions (REDG.E.ADD.F64) that resolve in L2 because consecutive blocks work on one patch.
// Floating-point contract: the reference is itself built with FMA contraction, so results agree to
// <= 1e-12 of each array's max-abs (tests/test_gpu_parity.py), not bit-for-bit.
#include "lpic_common.cuh"

namespace {

struct PatchView {
    const double *ex, *ey, *ez, *bx, *by, *bz;
    double *jx, *jy, *jz, *rho;
    double x0, y0, z0;
};

__device__ __forceinline__ PatchView patch_view(const Geom &g, double *F, const double *x0, const double *y0,
                                                const double *z0, int p) {
    const size_t stride = (size_t)g.npatch * g.ncell;
    double *base = F + (size_t)p * g.ncell;
    PatchView v;
    v.ex = base + LPIC_EX * stride; v.ey = base + LPIC_EY * stride; v.ez = base + LPIC_EZ * stride;
    v.bx = base + LPIC_BX * stride; v.by = base + LPIC_BY * stride; v.bz = base + LPIC_BZ * stride;
    v.jx = base + LPIC_JX * stride; v.jy = base + LPIC_JY * stride; v.jz = base + LPIC_JZ * stride;
    v.rho = base + LPIC_RHO * stride;
    v.x0 = x0[p]; v.y0 = y0[p]; v.z0 = z0[p];
    return v;
}

struct Slots {  // SoA attribute arenas of one species
    double *x, *y, *z, *w, *ux, *uy, *uz, *ig;
    double *part[6];
    const u8 *dead;
    const i64 *off, *npart;
};

__device__ __forceinline__ void tsc3(double d, double *g) {  // get_gx, unified_pusher_3d.c:65-70
    const double d2 = d * d;
    g[0] = 0.5 * (0.25 + d2 + d);
    g[1] = 0.75 - d2;
    g[2] = 0.5 * (0.25 + d2 - d);
}

// 27-point weighted sum, nesting z(y(x)) as interp_field_safe_3d (unified_pusher_3d.c:79-106).
// ox/oy/oz: storage offsets of the three stencil points along each axis (already wrapped and scaled).
__device__ __forceinline__ double gather27(const double *__restrict__ F, const double *fx, const double *fy,
                                           const double *fz, const int *ox, const int *oy, const int *oz) {
    double az[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        double ay[3];
#pragma unroll
        for (int b = 0; b < 3; b++) {
            const int base = oz[c] + oy[b];
            ay[b] = fx[0] * __ldg(F + base + ox[0]) + fx[1] * __ldg(F + base + ox[1]) + fx[2] * __ldg(F + base + ox[2]);
        }
        az[c] = fy[0] * ay[0] + fy[1] * ay[1] + fy[2] * ay[2];
    }
    return fz[0] * az[0] + fz[1] * az[1] + fz[2] * az[2];
}

__device__ __forceinline__ double gather9(const double *__restrict__ F, const double *fx, const double *fy,
                                          const int *ox, const int *oy) {
    double a[3];
#pragma unroll
    for (int b = 0; b < 3; b++)
        a[b] = fx[0] * __ldg(F + oy[b] + ox[0]) + fx[1] * __ldg(F + oy[b] + ox[1]) + fx[2] * __ldg(F + oy[b] + ox[2]);
    return fy[0] * a[0] + fy[1] * a[1] + fy[2] * a[2];
}

__device__ __forceinline__ void offsets3(int i, int N, int scale, int *o) {
    o[0] = wrapneg(i - 1, N) * scale;
    o[1] = wrapneg(i, N) * scale;
    o[2] = wrapneg(i + 1, N) * scale;
}

// E and B at the particle: ex(h,g,g) ey(g,h,g) ez(g,g,h) bx(g,h,h) by(h,g,h) bz(h,h,g)  (unified_pusher_3d.c:190-195)
template <int DIM>
__device__ __forceinline__ void gather_eb(const Geom &g, const PatchView &v, double x, double y, double z, double *eb) {
    const double X = (x - v.x0) * (1.0 / g.dx), Y = (y - v.y0) * (1.0 / g.dy);
    const double fX = floor(X), fY = floor(Y), rX = floor(X + 0.5), rY = floor(Y + 0.5);
    double gx[3], gy[3], hx[3], hy[3];
    tsc3(rX - X, gx); tsc3(fX - X + 0.5, hx);
    tsc3(rY - Y, gy); tsc3(fY - Y + 0.5, hy);
    int ogx[3], ohx[3], ogy[3], ohy[3];
    offsets3((int)rX, g.NX, g.NY * g.NZ, ogx); offsets3((int)fX, g.NX, g.NY * g.NZ, ohx);
    offsets3((int)rY, g.NY, g.NZ, ogy); offsets3((int)fY, g.NY, g.NZ, ohy);
    if (DIM == 3) {
        const double Z = (z - v.z0) * (1.0 / g.dz);
        const double fZ = floor(Z), rZ = floor(Z + 0.5);
        double gz[3], hz[3];
        tsc3(rZ - Z, gz); tsc3(fZ - Z + 0.5, hz);
        int ogz[3], ohz[3];
        offsets3((int)rZ, g.NZ, 1, ogz); offsets3((int)fZ, g.NZ, 1, ohz);
        eb[0] = gather27(v.ex, hx, gy, gz, ohx, ogy, ogz);
        eb[1] = gather27(v.ey, gx, hy, gz, ogx, ohy, ogz);
        eb[2] = gather27(v.ez, gx, gy, hz, ogx, ogy, ohz);
        eb[3] = gather27(v.bx, gx, hy, hz, ogx, ohy, ohz);
        eb[4] = gather27(v.by, hx, gy, hz, ohx, ogy, ohz);
        eb[5] = gather27(v.bz, hx, hy, gz, ohx, ohy, ogz);
    } else {
        eb[0] = gather9(v.ex, hx, gy, ohx, ogy);
        eb[1] = gather9(v.ey, gx, hy, ogx, ohy);
        eb[2] = gather9(v.ez, gx, gy, ogx, ogy);
        eb[3] = gather9(v.bx, gx, hy, ogx, ohy);
        eb[4] = gather9(v.by, hx, gy, ohx, ogy);
        eb[5] = gather9(v.bz, hx, hy, ohx, ohy);
    }
}

// unified_pusher_3d.c:15-51
__device__ __forceinline__ void boris_kick(double &ux, double &uy, double &uz, double &ig, const double *eb,
                                           double efactor, double bfactor) {
    const double umx = ux + efactor * eb[0], umy = uy + efactor * eb[1], umz = uz + efactor * eb[2];
    const double igm = 1.0 / sqrt(1.0 + umx * umx + umy * umy + umz * umz);
    const double Tx = bfactor * eb[3] * igm, Ty = bfactor * eb[4] * igm, Tz = bfactor * eb[5] * igm;
    const double upx = umx + umy * Tz - umz * Ty;
    const double upy = umy + umz * Tx - umx * Tz;
    const double upz = umz + umx * Ty - umy * Tx;
    const double Tf = 2.0 / (1.0 + Tx * Tx + Ty * Ty + Tz * Tz);
    const double Sx = Tf * Tx, Sy = Tf * Ty, Sz = Tf * Tz;
    ux = umx + upy * Sz - upz * Sy + efactor * eb[0];
    uy = umy + upz * Sx - upx * Sz + efactor * eb[1];
    uz = umz + upx * Sy - upy * Sx + efactor * eb[2];
    ig = 1.0 / sqrt(1.0 + ux * ux + uy * uy + uz * uz);
}

// calculate_S0 / calculate_S (current_deposit.h:7-35): 5-point arrays, `shift` in {-1,0,1} moves the 3 non-zeros.
__device__ __forceinline__ void shape5(double delta, int shift, double *S) {
    const double d2 = delta * delta;
    const double lo = 0.5 * (d2 + delta + 0.25), mid = 0.75 - d2, hi = 0.5 * (d2 - delta + 0.25);
    S[0] = shift < 0 ? lo : 0.0;
    S[1] = shift < 0 ? mid : (shift == 0 ? lo : 0.0);
    S[2] = shift < 0 ? hi : (shift == 0 ? mid : lo);
    S[3] = shift < 0 ? 0.0 : (shift == 0 ? hi : mid);
    S[4] = shift > 0 ? hi : 0.0;
}

__device__ __forceinline__ int wrap_base(int i, int N) {  // current_deposit.h:417-423
    i %= N;
    return i < 0 ? i + N : i;
}
__device__ __forceinline__ int wrap_once(int i, int N) { return i < 0 ? i + N : (i >= N ? i - N : i); }

struct DepositCoef3 {
    double q_dV, q_dydzdt, q_dxdzdt, q_dxdydt, dt;
};

// current_deposit_3d_fast + _cells (current_deposit.h:275-440).  x,y,z are the END-of-step positions; the
// deposit reconstructs +-dt/2 around them.  Loops are fully unrolled so every array index is static.
__device__ __forceinline__ void deposit3(const Geom &g, const PatchView &v, const DepositCoef3 &k, double x, double y,
                                         double z, double ux, double uy, double uz, double ig, double w) {
    const double vx = ux * LPIC_C_LIGHT * ig, vy = uy * LPIC_C_LIGHT * ig, vz = uz * LPIC_C_LIGHT * ig;
    const double X0 = (x - vx * 0.5 * k.dt - v.x0) / g.dx, X1 = (x + vx * 0.5 * k.dt - v.x0) / g.dx;
    const double Y0 = (y - vy * 0.5 * k.dt - v.y0) / g.dy, Y1 = (y + vy * 0.5 * k.dt - v.y0) / g.dy;
    const double Z0 = (z - vz * 0.5 * k.dt - v.z0) / g.dz, Z1 = (z + vz * 0.5 * k.dt - v.z0) / g.dz;
    const int ix0 = (int)floor(X0 + 0.5), iy0 = (int)floor(Y0 + 0.5), iz0 = (int)floor(Z0 + 0.5);
    const int ix1 = (int)floor(X1 + 0.5), iy1 = (int)floor(Y1 + 0.5), iz1 = (int)floor(Z1 + 0.5);
    const int dcx = ix1 - ix0, dcy = iy1 - iy0, dcz = iz1 - iz0;
    double S0x[5], S0y[5], S0z[5], S1x[5], S1y[5], S1z[5], DSx[5], DSy[5], DSz[5];
    shape5(ix0 - X0, 0, S0x); shape5(iy0 - Y0, 0, S0y); shape5(iz0 - Z0, 0, S0z);
    shape5(ix1 - X1, dcx, S1x); shape5(iy1 - Y1, dcy, S1y); shape5(iz1 - Z1, dcz, S1z);
#pragma unroll
    for (int i = 0; i < 5; i++) { DSx[i] = S1x[i] - S0x[i]; DSy[i] = S1y[i] - S0y[i]; DSz[i] = S1z[i] - S0z[i]; }
    const double cd = k.q_dV * w, fdx = k.q_dydzdt * w, fdy = k.q_dxdzdt * w, fdz = k.q_dxdydt * w;
    const int is = dcx < 0 ? 0 : 1, ie = dcx > 0 ? 5 : 4, js = dcy < 0 ? 0 : 1, je = dcy > 0 ? 5 : 4;
    const int ks = dcz < 0 ? 0 : 1, ke = dcz > 0 ? 5 : 4;
    const int bx0 = wrap_base(ix0, g.NX), by0 = wrap_base(iy0, g.NY), bz0 = wrap_base(iz0, g.NZ);
    double jxb[5][5];
#pragma unroll
    for (int a = 0; a < 5; a++)
#pragma unroll
        for (int b = 0; b < 5; b++) jxb[a][b] = 0.0;
#pragma unroll
    for (int i = 0; i < 5; i++) {
        if (i < is || i >= ie) continue;
        const int ox = wrap_once(bx0 + i - 2, g.NX) * g.NY * g.NZ;
        const double ax = S0x[i] + 0.5 * DSx[i], cx = 0.5 * S0x[i] + LPIC_ONE_THIRD * DSx[i], fx = fdx * DSx[i];
        double jyb[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int j = 0; j < 5; j++) {
            if (j < js || j >= je) continue;
            const int oy = ox + wrap_once(by0 + j - 2, g.NY) * g.NZ;
            const double ay = S0y[j] + 0.5 * DSy[j], cy = 0.5 * S0y[j] + LPIC_ONE_THIRD * DSy[j], fy = fdy * DSy[j];
            const double tz = ax * S0y[j] + cx * DSy[j];
            const double rxy = cd * S1x[i] * S1y[j];
            double jzb = 0.0;
#pragma unroll
            for (int kk = 0; kk < 5; kk++) {
                if (kk < ks || kk >= ke) continue;
                const int id = oy + wrap_once(bz0 + kk - 2, g.NZ);
                jxb[kk][j] -= fx * (ay * S0z[kk] + cy * DSz[kk]);
                jyb[kk] -= fy * (ax * S0z[kk] + cx * DSz[kk]);
                jzb -= fdz * DSz[kk] * tz;
                atomicAdd(v.jx + id, jxb[kk][j]);
                atomicAdd(v.jy + id, jyb[kk]);
                atomicAdd(v.jz + id, jzb);
                atomicAdd(v.rho + id, rxy * S1z[kk]);
            }
        }
    }
}

struct DepositCoef2 {
    double q_dxdy, q_dydt, q_dxdt, dt;
};

// current_deposit_2d_fast + _cells (current_deposit.h:150-268)
__device__ __forceinline__ void deposit2(const Geom &g, const PatchView &v, const DepositCoef2 &k, double x, double y,
                                         double ux, double uy, double uz, double ig, double w) {
    const double vx = ux * LPIC_C_LIGHT * ig, vy = uy * LPIC_C_LIGHT * ig, vz = uz * LPIC_C_LIGHT * ig;
    const double X0 = (x - vx * 0.5 * k.dt - v.x0) / g.dx, X1 = (x + vx * 0.5 * k.dt - v.x0) / g.dx;
    const double Y0 = (y - vy * 0.5 * k.dt - v.y0) / g.dy, Y1 = (y + vy * 0.5 * k.dt - v.y0) / g.dy;
    const int ix0 = (int)floor(X0 + 0.5), iy0 = (int)floor(Y0 + 0.5);
    const int ix1 = (int)floor(X1 + 0.5), iy1 = (int)floor(Y1 + 0.5);
    const int dcx = ix1 - ix0, dcy = iy1 - iy0;
    double S0x[5], S0y[5], S1x[5], S1y[5], DSx[5], DSy[5];
    shape5(ix0 - X0, 0, S0x); shape5(iy0 - Y0, 0, S0y);
    shape5(ix1 - X1, dcx, S1x); shape5(iy1 - Y1, dcy, S1y);
#pragma unroll
    for (int i = 0; i < 5; i++) { DSx[i] = S1x[i] - S0x[i]; DSy[i] = S1y[i] - S0y[i]; }
    const double cd = k.q_dxdy * w, fdx = k.q_dydt * w, fdy = k.q_dxdt * w, fvz = cd * vz;
    const double one_twelfth = 1.0 / 12.0;
    const int is = dcx < 0 ? 0 : 1, ie = dcx > 0 ? 5 : 4, js = dcy < 0 ? 0 : 1, je = dcy > 0 ? 5 : 4;
    double jxb[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int i = 0; i < 5; i++) {
        if (i < is || i >= ie) continue;
        const int ox = wrap_base(ix0 + i - 2, g.NX) * g.NY;
        const double a = S0x[i] + 0.5 * DSx[i], fxi = fdx * DSx[i], t12 = one_twelfth * DSx[i];
        double jyb = 0.0;
#pragma unroll
        for (int j = 0; j < 5; j++) {
            if (j < js || j >= je) continue;
            const int id = ox + wrap_base(iy0 + j - 2, g.NY);
            const double b = S0y[j] + 0.5 * DSy[j];
            jxb[j] -= fxi * b;
            jyb -= fdy * (DSy[j] * a);
            atomicAdd(v.jx + id, jxb[j]);
            atomicAdd(v.jy + id, jyb);
            atomicAdd(v.jz + id, fvz * (a * b + t12 * DSy[j]));
            atomicAdd(v.rho + id, cd * S1x[i] * S1y[j]);
        }
    }
}

// ---- block -> (patch, slot) mapping: blocks never straddle a patch ------------------------------------------
__device__ __forceinline__ bool my_slot(const Slots &s, int blocks_per_patch, int &p, i64 &slot) {
    p = blockIdx.x / blocks_per_patch;
    const i64 local = (i64)(blockIdx.x - p * blocks_per_patch) * blockDim.x + threadIdx.x;
    if (local >= s.npart[p]) return false;
    slot = s.off[p] + local;
    return true;
}

enum { MODE_FUSED = 0, MODE_GATHER, MODE_BORIS, MODE_POSITION, MODE_DEPOSIT };

// One species: half push, gather, Boris, half push, deposit (unified_pusher_3d.c:281-431).  Dead or
// NaN-position slots are skipped (:336-339, 387-388).
template <int DIM, int MODE, bool WRITE_PART>
__global__ void __launch_bounds__(128) k_particles(Geom g, double *__restrict__ F, const double *__restrict__ px0,
                                                   const double *__restrict__ py0, const double *__restrict__ pz0,
                                                   Slots s, int blocks_per_patch, double dt, double q, double m) {
    int p;
    i64 ip;
    if (!my_slot(s, blocks_per_patch, p, ip)) return;
    if (s.dead[ip]) return;
    double x = s.x[ip], y = s.y[ip], z = DIM == 3 ? s.z[ip] : 0.0;
    if (MODE != MODE_BORIS && (isnan(x) || isnan(y) || (DIM == 3 && isnan(z)))) return;
    const PatchView v = patch_view(g, F, px0, py0, pz0, p);
    double ux = s.ux[ip], uy = s.uy[ip], uz = s.uz[ip], ig = s.ig[ip];
    const double cdt = LPIC_C_LIGHT * 0.5 * dt;
    if (MODE == MODE_FUSED || MODE == MODE_GATHER || MODE == MODE_BORIS) {
        double eb[6];
        if (MODE == MODE_FUSED) {
            x += cdt * ig * ux;
            y += cdt * ig * uy;
            if (DIM == 3) z += cdt * ig * uz;
        }
        if (MODE == MODE_BORIS) {
#pragma unroll
            for (int a = 0; a < 6; a++) eb[a] = s.part[a][ip];
        } else {
            gather_eb<DIM>(g, v, x, y, z, eb);
            if (WRITE_PART || MODE == MODE_GATHER) {
#pragma unroll
                for (int a = 0; a < 6; a++) s.part[a][ip] = eb[a];
            }
        }
        if (MODE == MODE_GATHER) return;
        const double efactor = q * dt / (2 * m * LPIC_C_LIGHT), bfactor = q * dt / (2 * m);
        boris_kick(ux, uy, uz, ig, eb, efactor, bfactor);
        s.ux[ip] = ux; s.uy[ip] = uy; s.uz[ip] = uz; s.ig[ip] = ig;
        if (MODE == MODE_BORIS) return;
        x += cdt * ig * ux;
        y += cdt * ig * uy;
        s.x[ip] = x; s.y[ip] = y;
        if (DIM == 3) { z += cdt * ig * uz; s.z[ip] = z; }
    }
    if (MODE == MODE_POSITION) {  // push_position_2d (core/pusher/cpu.py:58-70): x += c*dt*inv_gamma*u
        s.x[ip] = x + LPIC_C_LIGHT * dt * ig * ux;
        s.y[ip] = y + LPIC_C_LIGHT * dt * ig * uy;
        if (DIM == 3) s.z[ip] = z + LPIC_C_LIGHT * dt * ig * uz;
        return;
    }
    const double w = s.w[ip];
    if (DIM == 3) {
        DepositCoef3 k;
        k.q_dV = q / (g.dx * g.dy * g.dz); k.q_dydzdt = q / (g.dy * g.dz * dt);
        k.q_dxdzdt = q / (g.dx * g.dz * dt); k.q_dxdydt = q / (g.dx * g.dy * dt); k.dt = dt;
        deposit3(g, v, k, x, y, z, ux, uy, uz, ig, w);
    } else {
        DepositCoef2 k;
        k.q_dxdy = q / (g.dx * g.dy); k.q_dydt = q / (g.dy * dt); k.q_dxdt = q / (g.dx * dt); k.dt = dt;
        deposit2(g, v, k, x, y, ux, uy, uz, ig, w);
    }
}

// ---- reductions -------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// which: 0 -> out[0]+=sum w, out[1]+=sum w*ux ; 1 -> out[0]+=sum w*(gamma-1) ; 2 -> cnt[0]+=alive
__global__ void __launch_bounds__(256) k_reduce(Slots s, int blocks_per_patch, int which, double *out, unsigned long long *cnt) {
    int p;
    i64 ip;
    double a = 0.0, b = 0.0;
    unsigned n = 0;
    if (my_slot(s, blocks_per_patch, p, ip) && !s.dead[ip]) {
        n = 1;
        if (which == 0) { a = s.w[ip]; b = a * s.ux[ip]; }
        if (which == 1) {
            const double u2 = s.ux[ip] * s.ux[ip] + s.uy[ip] * s.uy[ip] + s.uz[ip] * s.uz[ip];
            a = s.w[ip] * (u2 / (1.0 + sqrt(1.0 + u2)));  // gamma - 1 without cancellation
        }
    }
    a = warp_sum(a);
    b = warp_sum(b);
    n = __reduce_add_sync(0xffffffffu, n);
    if ((threadIdx.x & 31) == 0) {
        if (which == 2) { if (n) atomicAdd(cnt, (unsigned long long)n); }
        else { atomicAdd(out, a); if (which == 0) atomicAdd(out + 1, b); }
    }
}

// ---- synthetic loader -------------------------------------------------------------------------------------
__device__ __forceinline__ u64 mix64(u64 z) {  // splitmix64 finaliser
    z += 0x9e3779b97f4a7c15ull;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ double u01(u64 h) { return ((h >> 11) + 0.5) * (1.0 / 9007199254740992.0); }

// ppc particles per node, uniform in [x_i - dx/2, x_i + dx/2) (core/patch/cpu.py:88-99), cell-major order
// i, j, k as the reference loader; momenta ~ N(0, uth) per component (Box-Muller).
__global__ void __launch_bounds__(256) k_init_uniform(Geom g, const double *px0, const double *py0, const double *pz0,
                                                      Slots s, u8 *dead, double *id, int blocks_per_patch, i64 ppc,
                                                      double weight, double uth, u64 seed, u64 rank,
                                                      const i64 *patch_index) {
    int p;
    i64 ip;
    if (!my_slot(s, blocks_per_patch, p, ip)) return;
    const i64 local = ip - s.off[p];
    const i64 ncell = (i64)g.nx * g.ny * g.nz;
    const u64 bits = (rank << 50) | ((u64)patch_index[p] << 32) | (u64)local;
    id[ip] = __longlong_as_double((long long)bits);
    if (local >= ncell * ppc) {  // spare capacity: dead slot as ParticlesBase.extend leaves it
        const double nan = __longlong_as_double(0x7ff8000000000000ll);
        s.x[ip] = nan; s.y[ip] = nan; if (g.dim == 3) s.z[ip] = nan;
        s.ux[ip] = nan; s.uy[ip] = nan; s.uz[ip] = nan; s.ig[ip] = nan; s.w[ip] = 0.0;
        dead[ip] = 1;
        return;
    }
    const i64 cell = local / ppc;
    const int k = (int)(cell % g.nz), j = (int)((cell / g.nz) % g.ny), i = (int)(cell / ((i64)g.nz * g.ny));
    u64 h = mix64(seed ^ mix64(((u64)patch_index[p] << 40) ^ (u64)local));
    const double r0 = u01(h); h = mix64(h);
    const double r1 = u01(h); h = mix64(h);
    const double r2 = u01(h); h = mix64(h);
    s.x[ip] = px0[p] + (i + r0 - 0.5) * g.dx;
    s.y[ip] = py0[p] + (j + r1 - 0.5) * g.dy;
    if (g.dim == 3) s.z[ip] = pz0[p] + (k + r2 - 0.5) * g.dz;
    const double a0 = u01(h); h = mix64(h);
    const double a1 = u01(h); h = mix64(h);
    const double a2 = u01(h); h = mix64(h);
    const double a3 = u01(h);
    const double m0 = sqrt(-2.0 * log(a0)), m1 = sqrt(-2.0 * log(a2));
    const double ux = uth * m0 * cospi(2.0 * a1), uy = uth * m0 * sinpi(2.0 * a1), uz = uth * m1 * cospi(2.0 * a3);
    s.ux[ip] = ux; s.uy[ip] = uy; s.uz[ip] = uz;
    s.ig[ip] = 1.0 / sqrt(1.0 + ux * ux + uy * uy + uz * uz);
    s.w[ip] = weight;
    dead[ip] = 0;
}

Slots make_slots(const Species &sp) {
    Slots s;
    s.x = sp.attr[LPIC_P_X]; s.y = sp.attr[LPIC_P_Y]; s.z = sp.attr[LPIC_P_Z]; s.w = sp.attr[LPIC_P_W];
    s.ux = sp.attr[LPIC_P_UX]; s.uy = sp.attr[LPIC_P_UY]; s.uz = sp.attr[LPIC_P_UZ]; s.ig = sp.attr[LPIC_P_INV_GAMMA];
    for (int a = 0; a < 6; a++) s.part[a] = sp.attr[LPIC_P_EX_PART + a];
    s.dead = sp.dead; s.off = sp.d_off; s.npart = sp.d_npart;
    return s;
}

template <int MODE>
int launch_particles(lpic_ctx *c, int ispec, double dt, double q, double m, bool write_part) {
    REQUIRE(ispec >= 0 && ispec < c->nspec && c->spec[ispec].allocated, "species %d not allocated", ispec);
    Species &sp = c->spec[ispec];
    if ((write_part || MODE == MODE_GATHER || MODE == MODE_BORIS) && !sp.with_part) {
        lpic_set_error("species %d was allocated without ex_part..bz_part", ispec);
        return -2;
    }
    if (sp.max_npart == 0) return 0;
    const Geom &g = c->g;
    const int B = 128;
    const int bpp = (int)div_up(sp.max_npart, B);
    const unsigned grid = (unsigned)((i64)bpp * g.npatch);
    Slots s = make_slots(sp);
#define LAUNCH(D, W) k_particles<D, MODE, W><<<grid, B, 0, c->stream>>>(g, c->fields, c->d_x0, c->d_y0, c->d_z0, s, bpp, dt, q, m)
    if (g.dim == 3) { if (write_part) LAUNCH(3, true); else LAUNCH(3, false); }
    else { if (write_part) LAUNCH(2, true); else LAUNCH(2, false); }
#undef LAUNCH
    LAUNCHED(1);
    KERNEL_CHECK();
    return 0;
}

}  // namespace

extern "C" int lpic_push_deposit(lpic_ctx *c, int ispec, double dt, double q, double m, int flags) {
    return launch_particles<MODE_FUSED>(c, ispec, dt, q, m, (flags & LPIC_PUSH_WRITE_PART) != 0);
}
extern "C" int lpic_interpolate(lpic_ctx *c, int ispec) { return launch_particles<MODE_GATHER>(c, ispec, 0.0, 0.0, 1.0, false); }
extern "C" int lpic_push_momentum(lpic_ctx *c, int ispec, double dt, double q, double m) {
    return launch_particles<MODE_BORIS>(c, ispec, dt, q, m, false);
}
extern "C" int lpic_push_position(lpic_ctx *c, int ispec, double dt) {
    return launch_particles<MODE_POSITION>(c, ispec, dt, 0.0, 1.0, false);
}
extern "C" int lpic_deposit(lpic_ctx *c, int ispec, double dt, double q) {
    return launch_particles<MODE_DEPOSIT>(c, ispec, dt, q, 1.0, false);
}

static int reduce_species(lpic_ctx *c, int ispec, int which, double *outf, int nf, i64 *outi) {
    REQUIRE(ispec >= 0 && ispec < c->nspec && c->spec[ispec].allocated, "species %d not allocated", ispec);
    Species &sp = c->spec[ispec];
    CUDA_TRY(cudaMemsetAsync(c->d_tmpf, 0, 2 * sizeof(double), c->stream));
    CUDA_TRY(cudaMemsetAsync(c->d_tmp64, 0, sizeof(i64), c->stream));
    if (sp.max_npart > 0) {
        const int bpp = (int)div_up(sp.max_npart, 256);
        k_reduce<<<(unsigned)((i64)bpp * c->g.npatch), 256, 0, c->stream>>>(make_slots(sp), bpp, which, c->d_tmpf,
                                                                           (unsigned long long *)c->d_tmp64);
        LAUNCHED(1);
        KERNEL_CHECK();
    }
    if (outf) CUDA_TRY(cudaMemcpyAsync(outf, c->d_tmpf, nf * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if (outi) CUDA_TRY(cudaMemcpyAsync(outi, c->d_tmp64, sizeof(i64), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return 0;
}
extern "C" int lpic_weighted_drift(lpic_ctx *c, int ispec, double *out2) { return reduce_species(c, ispec, 0, out2, 2, nullptr); }
extern "C" int lpic_kinetic_sum(lpic_ctx *c, int ispec, double *out) { return reduce_species(c, ispec, 1, out, 1, nullptr); }
extern "C" int lpic_count_alive(lpic_ctx *c, int ispec, int64_t *out) { return reduce_species(c, ispec, 2, nullptr, 0, out); }

extern "C" int lpic_species_init_uniform(lpic_ctx *c, int ispec, int64_t ppc, double weight, double uth, uint64_t seed) {
    REQUIRE(ispec >= 0 && ispec < c->nspec && c->spec[ispec].allocated, "species %d not allocated", ispec);
    Species &sp = c->spec[ispec];
    const Geom &g = c->g;
    const i64 need = (i64)g.nx * g.ny * g.nz * ppc;
    for (int p = 0; p < g.npatch; p++) REQUIRE(sp.h_npart[p] >= need, "patch %d capacity %lld < %lld", p, (long long)sp.h_npart[p], (long long)need);
    i64 *d_pidx = c->d_tmp64 + 16;
    CUDA_TRY(cudaMemcpyAsync(d_pidx, c->h_patch_index, sizeof(i64) * g.npatch, cudaMemcpyHostToDevice, c->stream));
    const int bpp = (int)div_up(sp.max_npart, 256);
    k_init_uniform<<<(unsigned)((i64)bpp * g.npatch), 256, 0, c->stream>>>(g, c->d_x0, c->d_y0, c->d_z0, make_slots(sp), sp.dead,
                                                                         sp.attr[LPIC_P_ID], bpp, ppc, weight, uth, seed,
                                                                         (u64)c->rank, d_pidx);
    LAUNCHED(1);
    KERNEL_CHECK();
    sp.sort.valid = false;
    return 0;
}
