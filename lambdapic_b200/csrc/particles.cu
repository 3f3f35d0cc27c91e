// Field gather (quadratic spline on the staggered Yee grid) + relativistic Boris push + Esirkepov deposit.
//
// Reference behaviour restated (not copied): core/pusher/unified/unified_pusher_3d.c:15-217,281-431 and
// unified_pusher_2d.c:64-365 (half push, gather, Boris, half push), core/current/current_deposit.h:7-35,
// 150-268 (2D), 275-440 (3D) (charge-conserving deposit), core/interpolation/cpu3d.c:51-97,
// core/pusher/boris.py:6-38, core/pusher/cpu.py:58-90.
//
// Kernel structure (round 1): one thread per particle slot, blocks never straddle a patch so the patch's
// grid base pointers are block-uniform; the attribute streams (records or separate arrays, indexed through Slots::ps) are
// read/written in slot order; E/B are read
// through L1 (a patch tile + guards is 85 KB per component at 16^3); J/rho are accumulated with native
// fp64 global reductions (REDG.E.ADD.F64) that resolve in L2 because consecutive blocks work on one patch.
// Floating-point contract: the reference is itself built with FMA contraction, so results agree to
// <= 1e-12 of each array's max-abs (tests/test_gpu_parity.py), not bit-for-bit.
#include <stdlib.h>
#include <algorithm>
#include "lpic_common.cuh"
#include "particle_math.cuh"

namespace {

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
    double x = s.x[ip * s.ps], y = s.y[ip * s.ps], z = DIM == 3 ? s.z[ip * s.ps] : 0.0;
    if (MODE != MODE_BORIS && (isnan(x) || isnan(y) || (DIM == 3 && isnan(z)))) return;
    const PatchView v = patch_view(g, F, px0, py0, pz0, p);
    double ux = s.ux[ip * s.ps], uy = s.uy[ip * s.ps], uz = s.uz[ip * s.ps], ig = s.ig[ip * s.ps];
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
        s.ux[ip * s.ps] = ux; s.uy[ip * s.ps] = uy; s.uz[ip * s.ps] = uz; s.ig[ip * s.ps] = ig;
        if (MODE == MODE_BORIS) return;
        x += cdt * ig * ux;
        y += cdt * ig * uy;
        s.x[ip * s.ps] = x; s.y[ip * s.ps] = y;
        if (DIM == 3) { z += cdt * ig * uz; s.z[ip * s.ps] = z; }
    }
    if (MODE == MODE_POSITION) {  // push_position_2d (core/pusher/cpu.py:58-70): x += c*dt*inv_gamma*u
        s.x[ip * s.ps] = x + LPIC_C_LIGHT * dt * ig * ux;
        s.y[ip * s.ps] = y + LPIC_C_LIGHT * dt * ig * uy;
        if (DIM == 3) s.z[ip * s.ps] = z + LPIC_C_LIGHT * dt * ig * uz;
        return;
    }
    const double w = s.w[ip * s.ps];
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
// Persistent grid: CTAs stride over the patches, threads over the slots, partial sums stay in registers and every CTA
// issues ONE atomic per result at the end (one atomic per warp on a single address serialised in L2: 5.8 ms per species
// at 128^3 x 32 ppc, four times the kernel's memory time).
__global__ void __launch_bounds__(256) k_reduce(Slots s, int npatch, int which, double *out, unsigned long long *cnt) {
    __shared__ double sa[8], sb[8];
    __shared__ unsigned long long sn[8];
    double a = 0.0, b = 0.0;
    unsigned long long n = 0;
    for (int p = blockIdx.x; p < npatch; p += gridDim.x) {
        const i64 off = s.off[p];
        const int np = (int)s.npart[p];
        for (int t = threadIdx.x; t < np; t += blockDim.x) {
            const i64 ip = off + t;
            if (s.dead[ip]) continue;
            n++;
            if (which == 2) continue;
            double w, ux, uy, uz;
            if (s.ps == LPIC_NREC) {
                const double2 *r = reinterpret_cast<const double2 *>(s.x + ip * LPIC_NREC);
                const double2 r1 = r[1], r2 = r[2];
                w = r1.y; ux = r2.x; uy = r2.y; uz = which == 1 ? r[3].x : 0.0;
            } else {
                w = s.w[ip]; ux = s.ux[ip];
                uy = which == 1 ? s.uy[ip] : 0.0; uz = which == 1 ? s.uz[ip] : 0.0;
            }
            if (which == 0) { a += w; b += w * ux; }
            else {
                const double u2 = ux * ux + uy * uy + uz * uz;
                a += w * (u2 / (1.0 + sqrt(1.0 + u2)));  // gamma - 1 without cancellation
            }
        }
    }
    a = warp_sum(a);
    b = warp_sum(b);
    for (int o = 16; o; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) { sa[w] = a; sb[w] = b; sn[w] = n; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < 8; i++) { a += sa[i]; b += sb[i]; n += sn[i]; }
        if (which == 2) { if (n) atomicAdd(cnt, n); }
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
        s.x[ip * s.ps] = nan; s.y[ip * s.ps] = nan; s.z[ip * s.ps] = g.dim == 3 ? nan : 0.0;
        s.ux[ip * s.ps] = nan; s.uy[ip * s.ps] = nan; s.uz[ip * s.ps] = nan; s.ig[ip * s.ps] = nan; s.w[ip * s.ps] = 0.0;
        dead[ip] = 1;
        return;
    }
    const i64 cell = local / ppc;
    const int k = (int)(cell % g.nz), j = (int)((cell / g.nz) % g.ny), i = (int)(cell / ((i64)g.nz * g.ny));
    u64 h = mix64(seed ^ mix64(((u64)patch_index[p] << 40) ^ (u64)local));
    const double r0 = u01(h); h = mix64(h);
    const double r1 = u01(h); h = mix64(h);
    const double r2 = u01(h); h = mix64(h);
    s.x[ip * s.ps] = px0[p] + (i + r0 - 0.5) * g.dx;
    s.y[ip * s.ps] = py0[p] + (j + r1 - 0.5) * g.dy;
    s.z[ip * s.ps] = g.dim == 3 ? pz0[p] + (k + r2 - 0.5) * g.dz : 0.0;
    const double a0 = u01(h); h = mix64(h);
    const double a1 = u01(h); h = mix64(h);
    const double a2 = u01(h); h = mix64(h);
    const double a3 = u01(h);
    const double m0 = sqrt(-2.0 * log(a0)), m1 = sqrt(-2.0 * log(a2));
    const double ux = uth * m0 * cospi(2.0 * a1), uy = uth * m0 * sinpi(2.0 * a1), uz = uth * m1 * cospi(2.0 * a3);
    s.ux[ip * s.ps] = ux; s.uy[ip * s.ps] = uy; s.uz[ip * s.ps] = uz;
    s.ig[ip * s.ps] = 1.0 / sqrt(1.0 + ux * ux + uy * uy + uz * uz);
    s.w[ip * s.ps] = weight;
    dead[ip] = 0;
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
    sp.lists_valid = false; sp.sort.keys_valid = sp.sort.keys_written = false;  // positions change: the migration lists no longer describe the slots
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

int lpic_push_deposit_sorted(lpic_ctx *c, int ispec, double dt, double q, double m, bool write_part);
int lpic_push_deposit_tiles(lpic_ctx *c, int ispec, double dt, double q, double m, bool write_part);

extern "C" int lpic_push_deposit(lpic_ctx *c, int ispec, double dt, double q, double m, int flags) {
    DeviceGuard dg(c);
    const bool write_part = (flags & LPIC_PUSH_WRITE_PART) != 0;
    REQUIRE(ispec >= 0 && ispec < c->nspec && c->spec[ispec].allocated, "species %d not allocated", ispec);
    if (write_part && !c->spec[ispec].with_part) {
        lpic_set_error("species %d was allocated without ex_part..bz_part", ispec);
        return -2;
    }
    c->spec[ispec].lists_valid = false; c->spec[ispec].sort.keys_valid = c->spec[ispec].sort.keys_written = false;
    if (!(flags & LPIC_PUSH_SLOT_ORDER)) {
        // default in 3D: the tile kernel (push_tile.cu); 2D, patches too large for its histogram and LPIC_PUSH_SORTED=1
        // (A/B runs): the round-1 cell-ordered kernel (push_sorted.cu)
        int r = getenv("LPIC_PUSH_SORTED") ? 1 : lpic_push_deposit_tiles(c, ispec, dt, q, m, write_part);
        if (r <= 0) return r;
        r = lpic_push_deposit_sorted(c, ispec, dt, q, m, write_part);
        if (r <= 0) return r;
    }
    return launch_particles<MODE_FUSED>(c, ispec, dt, q, m, write_part);
}
extern "C" int lpic_interpolate(lpic_ctx *c, int ispec) {
    DeviceGuard dg(c); return launch_particles<MODE_GATHER>(c, ispec, 0.0, 0.0, 1.0, false); }
extern "C" int lpic_push_momentum(lpic_ctx *c, int ispec, double dt, double q, double m) {
    DeviceGuard dg(c);
    return launch_particles<MODE_BORIS>(c, ispec, dt, q, m, false);
}
extern "C" int lpic_push_position(lpic_ctx *c, int ispec, double dt) {
    DeviceGuard dg(c);
    return launch_particles<MODE_POSITION>(c, ispec, dt, 0.0, 1.0, false);
}
extern "C" int lpic_deposit(lpic_ctx *c, int ispec, double dt, double q) {
    DeviceGuard dg(c);
    return launch_particles<MODE_DEPOSIT>(c, ispec, dt, q, 1.0, false);
}

static int reduce_species(lpic_ctx *c, int ispec, int which, double *outf, int nf, i64 *outi) {
    REQUIRE(ispec >= 0 && ispec < c->nspec && c->spec[ispec].allocated, "species %d not allocated", ispec);
    Species &sp = c->spec[ispec];
    CUDA_TRY(cudaMemsetAsync(c->d_tmpf, 0, 2 * sizeof(double), c->stream));
    CUDA_TRY(cudaMemsetAsync(c->d_tmp64, 0, sizeof(i64), c->stream));
    if (sp.max_npart > 0) {
        const unsigned grid = (unsigned)std::min<i64>(c->g.npatch, 148 * 8);
        k_reduce<<<grid, 256, 0, c->stream>>>(make_slots(sp), c->g.npatch, which, c->d_tmpf, (unsigned long long *)c->d_tmp64);
        LAUNCHED(1);
        KERNEL_CHECK();
    }
    if (outf) CUDA_TRY(cudaMemcpyAsync(outf, c->d_tmpf, nf * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if (outi) CUDA_TRY(cudaMemcpyAsync(outi, c->d_tmp64, sizeof(i64), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return 0;
}
extern "C" int lpic_weighted_drift(lpic_ctx *c, int ispec, double *out2) {
    DeviceGuard dg(c); return reduce_species(c, ispec, 0, out2, 2, nullptr); }
extern "C" int lpic_kinetic_sum(lpic_ctx *c, int ispec, double *out) {
    DeviceGuard dg(c); return reduce_species(c, ispec, 1, out, 1, nullptr); }
extern "C" int lpic_count_alive(lpic_ctx *c, int ispec, int64_t *out) {
    DeviceGuard dg(c); return reduce_species(c, ispec, 2, nullptr, 0, out); }

extern "C" int lpic_species_init_uniform(lpic_ctx *c, int ispec, int64_t ppc, double weight, double uth, uint64_t seed) {
    DeviceGuard dg(c);
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
    sp.lists_valid = false; sp.sort.keys_valid = sp.sort.keys_written = false;
    return 0;
}
