// Fused ConvNeXt MLP on tcgen05 for the C = 256, H = 1024 blocks of the vector estimator and the text encoder (sm_100a):
//
//     x[rows, C]  <-  ( x + gamma * ( GELU( a W1 + b1 ) W2 + b2 ) ) * mask          a = LayerNorm(dwconv(x)) as split-bf16
//
// Why: as two GEMMs the block is bound by operand ingest into the SMs — the 19 MB hidden tensor is written,
// then re-read once per N tile, and every tile re-reads its share of W. Here the hidden activations never leave the SM.
//
// One thread-block CLUSTER of 4 CTAs owns a 128-row tile; CTA c owns hidden units [256c, 256c+256):
//   phase 1  S_c[128 x 256]  = a[128 x 256] . W1[256c.., :]^T            (TMEM columns 0..255)
//   epi 1    P_c = split-bf16( GELU(S_c + b1) ) written as a K-major swizzled A operand into the smem the a-tile occupied
//   phase 2  O_c[128 x 256] = P_c[128 x 256] . W2[:, 256c..]^T           (TMEM columns 256..511)  — a partial sum over hidden units
//   reduce   CTA p owns output columns [64p, 64p+64): every CTA ships its partial for those columns into p's shared memory
//            over DSMEM (st.shared::cluster), p adds the four partials in rank order (deterministic) and applies
//            bias / layer-scale / residual / mask with coalesced global accesses.
// Per SM ingest: a-tile 128 KB + a quarter of W1 and of W2 (2 x 256 KB), against ~1.3 MB for the same rows as two GEMMs.
//
// Warps: 0 = TMA producer (a tile once, then 16 weight units of 128 rows x 64 K through a 3-slot ring), 1 = MMA issuer,
// 2..9 = epilogue (TMEM lane quarter = warp % 4, column half = (warp - 2) / 4). Arithmetic: split-bf16, 3 MMAs per K slice.
#pragma once
#include "gemm_tc.cuh"

namespace stc {
namespace mlp {

constexpr int C = 256, H = 1024, CS = 4, HC = H / CS;     // channels, hidden units, cluster size, hidden units per CTA
constexpr int BM = 128, BK = 64, UMMA_K = 16;
constexpr int NUM_THREADS = 320;
constexpr int KBLK = BM * BK * 2;                 // 16 KB: 128 rows x 64 bf16, one half (hi or lo)
constexpr int X_BYTES = 2 * (C / BK) * KBLK;      // 128 KB: a-tile (hi k-blocks 0..3, lo k-blocks 0..3); later P, later staging
constexpr int UNIT = 2 * KBLK;                    // 32 KB: 128 weight rows x 64 K, hi + lo
constexpr int SLOTS = 3;
constexpr int OFF_X = 0, OFF_RING = X_BYTES, OFF_BAR = OFF_RING + SLOTS * UNIT;
constexpr int OFF_B1 = OFF_BAR + 256;             // 1 KB: this CTA's slice of the pw1 bias (read by the GELU epilogue)
constexpr int SMEM_BYTES = OFF_B1 + 1024 + 1024;
static_assert(SMEM_BYTES <= 232448, "shared memory budget");
constexpr int UNITS_PER_PHASE = (C / BK) * (HC / 128);     // 8
constexpr int RECV_BYTES = BM * 64 * 4;           // 32 KB: one sender's partial for my 64 columns
static_assert(3 * RECV_BYTES <= SLOTS * UNIT, "receive slots alias the weight ring");
static_assert(HC == C, "phase 1 and phase 2 share the unit schedule");
constexpr int HC_THIN = 128, CS_THIN = H / HC_THIN;        // "thin" split form: eight hidden slices of 128 per row tile (small batches)
constexpr int HC_THIN64 = 64, CS_THIN64 = H / HC_THIN64;   // sixteen slices of 64 (<= 9 row tiles: the batch-1 latency path)

struct Params {
    int M;
    const float* b1; const float* b2; const float* gamma; const float* mask;
    float* x;                       // [M, C] residual stream, updated in place
    float* partial;                 // split variant: [CS][M rounded up to 128][C] fp32 partial outputs (reduced by mlp_reduce_kernel)
    // Producer mode (dw_wT != null; opt-in, measured slower — see Handle::mlp_producer): the a-tile is not loaded but COMPUTED
    // here — depthwise conv (+bias) -> LayerNorm of the residual stream x, straight into the swizzled operand layout.
    const float* dw_wT; const float* dw_b; const float* ln_g; const float* ln_b;     // taps [K][C], bias, LayerNorm scale / shift
    const int* off; int B; int K, dil, pad_left; float eps;                          // packed sequences of x, conv geometry
    long long* trace;               // debug (stc_debug_mlp with STC_MLP_TRACE=1): clock64() stamps of CTA 0's pipeline events
    // In-kernel reduce (split forms, cooperative launch with grid <= SM count so that every CTA is resident): cnt != null. After its
    // partial is in global memory a CTA arrives on its row tile's counter, waits for the tile's other hidden-slice CTAs and then
    // finishes BM / CSt rows of the tile itself — what mlp_reduce_kernel / mlp_reduce_post_kernel would do in a launch of their own.
    int* cnt;                       // [row tiles][2]: arrivals, departures (self-resetting)
    const float* add_vec; const float* post_ln_g; const float* post_ln_b;       // post-ops, see mlp_reduce_post_kernel
    __nv_bfloat16* out_hi; __nv_bfloat16* out_lo;
};
#define STC_TRACE(idx) do { if (p.trace && blockIdx.x == 0 && lane == 0) p.trace[idx] = clock64(); } while (0)

constexpr int KMAX = 5;             // depthwise taps supported by the producer mode

STC_DEVINL uint32_t mapa_u32(uint32_t local, uint32_t rank) {
    uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank)); return r;
}
STC_DEVINL void st_cluster_v4(uint32_t addr, float a, float b, float c, float d) {
    asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// kCluster: the four CTAs of a row tile form a cluster and reduce over DSMEM (needs all clusters resident in one wave: 33 fit
// on a B200). !kCluster: four independent CTAs (blockIdx.x % 4 = hidden slice) write their partial outputs to global scratch
// and mlp_reduce_kernel finishes the block — no placement constraint, 37 row tiles fill the 148 SMs.
// HCt: hidden units per CTA. 256 (CS = 4 CTAs per row tile) everywhere the SMs are full; 128 (the "thin" form, eight CTAs per row
// tile, !kCluster only) when 8 x row tiles still fit one wave: a CTA's serial chain a-tile -> S -> GELU -> O halves (18 -> ~12 us
// for a single row tile — the batch-1 latency path runs 172 of these per utterance).
// kWide (HCt = 256 only): the MMAs are N = 256 wide. A weight unit is then 256 rows x 64 K of ONE half — unit (kb, 0) the hi halves,
// unit (kb, 1) the lo halves, same 32 KB and the same ring — and the hi unit feeds a_lo.w_hi and a_hi.w_hi, the lo unit a_hi.w_lo.
// Why: an N = 128 MMA reads 8 KB of operands (4 KB of `a`, 4 KB of W) per 64 tensor cycles = 128 B/clk, all the shared-memory port
// has, while TMA is filling the ring through the same port (traces: 1150-1500 cycles per unit against 768 of pure MMA time); an
// N = 256 MMA reads 12 KB per 128 cycles = 96 B/clk.
template <bool kCluster, int HCt = HC, bool kWide = false>
STC_DEVINL void convnext_mlp_body(const CUtensorMap& map_a_hi, const CUtensorMap& map_a_lo, const CUtensorMap& map_w1_hi,
                                  const CUtensorMap& map_w1_lo, const CUtensorMap& map_w2_hi, const CUtensorMap& map_w2_lo,
                                  const Params& p) {
    using namespace tc;
    pdl_trigger();
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t bar = smem_base + OFF_BAR;
    const uint32_t bar_a = bar, bar_s1 = bar + 8, bar_o = bar + 16;
    auto full_bar = [&](int s) { return bar + 24 + 8u * s; };
    auto empty_bar = [&](int s) { return bar + 48 + 8u * s; };
    auto bar_p = [&](int j) { return bar + 72 + 8u * j; };
    // a-tile K block kb (hi + lo, 32 KB) has landed: the first MMAs start after 32 KB of `a` + one weight unit instead of after the
    // whole 128 KB tile (producer mode keeps the single bar_a)
    auto bar_ak = [&](int kb) { return kb == 0 ? bar_a : bar + 160 + 8u * kb; };
    const uint32_t tmem_slot = bar + 104;
    volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + OFF_BAR + 104);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    static_assert(HCt == HC || ((HCt == HC_THIN || HCt == HC_THIN64) && !kCluster), "hidden slice per CTA");
    static_assert(!kWide || (HCt == HC && !kCluster), "wide MMAs: the 256-unit split form");
    constexpr int CSt = H / HCt;
    constexpr int WR1 = HCt < 128 ? HCt : 128;                                    // weight rows of a phase-1 unit (= its MMA N)
    constexpr int U1 = (C / BK) * (HCt / WR1), U2 = (HCt / BK) * (C / 128);       // weight units of phase 1 / phase 2
    const int crank = kCluster ? (int)cluster_ctarank() : (int)(blockIdx.x % CSt);
    const int m0 = (kCluster ? (int)cluster_id_x() : (int)(blockIdx.x / CSt)) * BM;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a_hi); tma_prefetch_desc(&map_a_lo); tma_prefetch_desc(&map_w1_hi);
        tma_prefetch_desc(&map_w1_lo); tma_prefetch_desc(&map_w2_hi); tma_prefetch_desc(&map_w2_lo);
        mbar_init(bar_a, p.dw_wT ? 8 : 1); mbar_init(bar_s1, 1); mbar_init(bar_o, 1);
        for (int kb = 1; kb < C / BK; ++kb) mbar_init(bar_ak(kb), 1);
        for (int s = 0; s < SLOTS; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int j = 0; j < HCt / BK; ++j) mbar_init(bar_p(j), 8);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_gen;
    pdl_wait();

    if (warp == 0) {
        if (elect_one()) {
            auto load_a = [&](int kb) {
                mbar_expect_tx(bar_ak(kb), 2 * KBLK);
                tma_load_2d(smem_base + OFF_X + kb * KBLK, &map_a_hi, bar_ak(kb), kb * BK, m0);
                tma_load_2d(smem_base + OFF_X + (C / BK + kb) * KBLK, &map_a_lo, bar_ak(kb), kb * BK, m0);
            };
            if (!p.dw_wT) load_a(0);
            for (int u = 0; u < U1 + U2; ++u) {
                const int s = u % SLOTS;
                mbar_wait(empty_bar(s), ((u / SLOTS) & 1) ^ 1);
                const uint32_t dst = smem_base + OFF_RING + s * UNIT;
                const bool second = u >= U1;
                mbar_expect_tx(full_bar(s), second ? UNIT : 2 * WR1 * BK * 2);
                const int v = second ? u - U1 : u;
                const int kb = (second || HCt == HC) ? v >> 1 : v, nh = (second || HCt == HC) ? v & 1 : 0;
                if constexpr (kWide) {              // one half (nh: 0 = hi, 1 = lo) of 256 weight rows (map boxes of 256 rows)
                    if (!second) tma_load_2d(dst, nh ? &map_w1_lo : &map_w1_hi, full_bar(s), kb * BK, crank * HCt);
                    else tma_load_2d(dst, nh ? &map_w2_lo : &map_w2_hi, full_bar(s), crank * HCt + kb * BK, 0);
                } else if (!second) {               // W1[hidden rows, C]: rows crank*HCt + nh*128, K block kb of C
                    tma_load_2d(dst, &map_w1_hi, full_bar(s), kb * BK, crank * HCt + nh * 128);
                    tma_load_2d(dst + KBLK, &map_w1_lo, full_bar(s), kb * BK, crank * HCt + nh * 128);
                } else {                            // W2[C rows, hidden]: rows nh*128, K block kb of this CTA's hidden slice
                    tma_load_2d(dst, &map_w2_hi, full_bar(s), crank * HCt + kb * BK, nh * 128);
                    tma_load_2d(dst + KBLK, &map_w2_lo, full_bar(s), crank * HCt + kb * BK, nh * 128);
                }
                if (u == 0 && !p.dw_wT) {           // the rest of the a-tile queues behind the first weight unit
                    for (int k2 = 1; k2 < C / BK; ++k2) load_a(k2);
                }
            }
        }
        __syncwarp();                      // reconverge before the (warp-aligned) cluster barriers below
    } else if (warp == 1) {
        constexpr uint32_t idesc2 = make_idesc_bf16(BM, 128), idesc1 = make_idesc_bf16(BM, WR1);
        if (p.dw_wT) mbar_wait(bar_a, 0);
        for (int u = 0; u < U1 + U2; ++u) {
            const bool second = u >= U1;
            const int s = u % SLOTS, v = second ? u - U1 : u;
            const int kb = (second || HCt == HC) ? v >> 1 : v, nh = (second || HCt == HC) ? v & 1 : 0;
            if (!second && nh == 0 && !p.dw_wT) mbar_wait(bar_ak(kb), 0);       // a-tile K block kb has landed
            if (second && nh == 0) mbar_wait(bar_p(kb), 0);             // P k-block kb written by the epilogue warps
            mbar_wait(full_bar(s), (u / SLOTS) & 1);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t xk = smem_base + OFF_X + kb * KBLK, st = smem_base + OFF_RING + s * UNIT;
                const uint64_t a_hi = make_smem_desc(xk), a_lo = make_smem_desc(xk + (C / BK) * KBLK);
                const uint64_t w_hi = make_smem_desc(st), w_lo = make_smem_desc(st + KBLK);
                if constexpr (kWide) {
                    constexpr uint32_t idw = make_idesc_bf16(BM, 256);
                    const uint32_t dw = tmem_base + (second ? 256 : 0);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        const uint64_t adv = (uint64_t)((k * UMMA_K * 2) >> 4);
                        if (nh == 0) {              // w_hi unit
                            umma_bf16(dw, a_lo + adv, w_hi + adv, idw, (kb | k) != 0);
                            umma_bf16(dw, a_hi + adv, w_hi + adv, idw, 1);
                        } else umma_bf16(dw, a_hi + adv, w_hi + adv, idw, 1);       // the unit holds the lo halves
                    }
                } else {
                const uint32_t d = tmem_base + (second ? 256 : 0) + nh * 128;
                const uint32_t idesc = second ? idesc2 : idesc1;
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; ++k) {
                    const uint64_t adv = (uint64_t)((k * UMMA_K * 2) >> 4);
                    umma_bf16(d, a_lo + adv, w_hi + adv, idesc, (kb | k) != 0);
                    umma_bf16(d, a_hi + adv, w_lo + adv, idesc, 1);
                    umma_bf16(d, a_hi + adv, w_hi + adv, idesc, 1);
                }
                }
                umma_commit(empty_bar(s));
                if (u == U1 - 1) umma_commit(bar_s1);                   // S complete; the a-tile is dead
                if (u == U1 + U2 - 1) umma_commit(bar_o);               // partial O complete; ring and P are dead
            }
            __syncwarp();
        }
    } else {
        if (p.dw_wT) {
            // ===== a-tile producer: warp w takes rows w-2, w+6, ... two at a time (both rows' taps in flight together);
            //       lane owns channels [8 lane, 8 lane + 8) = one 16-byte chunk of k-block lane / 8 =====
            const int c0 = lane * 8;
            uint8_t* xa = smem_gen + OFF_X + (lane >> 3) * KBLK;
            int bseq = -2;
#pragma unroll 1
            for (int rl = warp - 2; rl < BM; rl += 16) {
                float4 xv[2][KMAX][2];
                bool ok[2][KMAX];
                int seq_ok[2];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int row = m0 + rl + 8 * u;
                    int b = -1;
                    if (row < p.M) {
                        if (bseq == -2 || bseq < 0) b = find_seq(p.off, p.B, row);
                        else { b = bseq; while (b < p.B && row >= __ldg(p.off + b + 1)) ++b; if (b >= p.B) b = -1; }
                        bseq = b;
                    }
                    seq_ok[u] = b;
                    const int base = b >= 0 ? __ldg(p.off + b) : 0, n = row - base, N = b >= 0 ? __ldg(p.off + b + 1) - base : 0;
#pragma unroll
                    for (int k = 0; k < KMAX; ++k) {
                        const int nn = n + k * p.dil - p.pad_left;
                        ok[u][k] = b >= 0 && k < p.K && nn >= 0 && nn < N;
                        const float* xr = p.x + ((size_t)base + (ok[u][k] ? nn : 0)) * C + c0;
                        if (ok[u][k]) { xv[u][k][0] = *reinterpret_cast<const float4*>(xr); xv[u][k][1] = *reinterpret_cast<const float4*>(xr + 4); }
                    }
                }
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int r = rl + 8 * u;
                    float y[8];
                    if (seq_ok[u] < 0) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) y[j] = 0.f;                      // rows past M / bucket padding: finite operand
                    } else {
                        const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.dw_b + c0)), b1v = __ldg(reinterpret_cast<const float4*>(p.dw_b + c0 + 4));
                        y[0] = b0.x; y[1] = b0.y; y[2] = b0.z; y[3] = b0.w; y[4] = b1v.x; y[5] = b1v.y; y[6] = b1v.z; y[7] = b1v.w;
#pragma unroll
                        for (int k = 0; k < KMAX; ++k) {
                            if (!ok[u][k]) continue;
                            const float4 w0 = __ldg(reinterpret_cast<const float4*>(p.dw_wT + (size_t)k * C + c0));
                            const float4 w1 = __ldg(reinterpret_cast<const float4*>(p.dw_wT + (size_t)k * C + c0 + 4));
                            y[0] += w0.x * xv[u][k][0].x; y[1] += w0.y * xv[u][k][0].y; y[2] += w0.z * xv[u][k][0].z; y[3] += w0.w * xv[u][k][0].w;
                            y[4] += w1.x * xv[u][k][1].x; y[5] += w1.y * xv[u][k][1].y; y[6] += w1.z * xv[u][k][1].z; y[7] += w1.w * xv[u][k][1].w;
                        }
                        float s = 0.f;
#pragma unroll
                        for (int j = 0; j < 8; ++j) s += y[j];
                        const float mean = warp_sum<float>(s) / (float)C;
                        float v = 0.f;
#pragma unroll
                        for (int j = 0; j < 8; ++j) { y[j] -= mean; v += y[j] * y[j]; }
                        const float den = sqrtf(warp_sum<float>(v) / (float)C + p.eps);
                        const float4 g0 = __ldg(reinterpret_cast<const float4*>(p.ln_g + c0)), g1 = __ldg(reinterpret_cast<const float4*>(p.ln_g + c0 + 4));
                        const float4 h0 = __ldg(reinterpret_cast<const float4*>(p.ln_b + c0)), h1 = __ldg(reinterpret_cast<const float4*>(p.ln_b + c0 + 4));
                        y[0] = y[0] / den * g0.x + h0.x; y[1] = y[1] / den * g0.y + h0.y; y[2] = y[2] / den * g0.z + h0.z; y[3] = y[3] / den * g0.w + h0.w;
                        y[4] = y[4] / den * g1.x + h1.x; y[5] = y[5] / den * g1.y + h1.y; y[6] = y[6] / den * g1.z + h1.z; y[7] = y[7] / den * g1.w + h1.w;
                    }
                    uint32_t hi[4], lo[4];
#pragma unroll
                    for (int t = 0; t < 4; ++t) split_pair(y[2 * t], y[2 * t + 1], hi[t], lo[t]);
                    uint8_t* dst = xa + (r >> 3) * 1024 + (r & 7) * 128 + (((lane & 7) ^ (r & 7)) * 16);
                    *reinterpret_cast<uint4*>(dst) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                    *reinterpret_cast<uint4*>(dst + (C / BK) * KBLK) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_a);
        }
        // ===== epilogue 1: P = split(GELU(S + b1)) =====
        // b1 slice -> shared memory while S is still being accumulated: with 226 KB of shared memory in use the SM has no L1, so a
        // __ldg inside the loop below would be an L2 round trip (~0.5 us under load) on the CTA's serial chain, once per K block
        float* b1s = reinterpret_cast<float*>(smem_gen + OFF_B1);
        {
            const int t = (int)threadIdx.x - 64;
            if (t < HCt) b1s[t] = __ldg(p.b1 + crank * HCt + t);
            asm volatile("bar.sync 1, 256;" ::: "memory");           // the eight epilogue warps only
        }
        const int q = warp & 3, half = (warp - 2) >> 2, r = q * 32 + lane;
        const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
        const uint32_t prow_off = (uint32_t)((r >> 3) * 1024 + (r & 7) * 128);
        mbar_wait(bar_s1, 0);
        tc_fence_after();
#pragma unroll 1
        for (int j = 0; j < HCt / BK; ++j) {
            uint32_t v[32];
            __syncwarp();
            tmem_ld32(trow + j * BK + half * 32, v);
            const float* b1 = b1s + j * BK + half * 32;
            uint8_t* p_hi = smem_gen + OFF_X + j * KBLK + prow_off;
            uint8_t* p_lo = p_hi + (C / BK) * KBLK;
#pragma unroll
            for (int c8 = 0; c8 < 4; ++c8) {
                uint32_t hi[4], lo[4];
                const float4 ba = *reinterpret_cast<const float4*>(b1 + c8 * 8), bb = *reinterpret_cast<const float4*>(b1 + c8 * 8 + 4);   // broadcast
                const float bv[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const float e0 = gelu_erf_mufu(__uint_as_float(v[c8 * 8 + 2 * t]) + bv[2 * t]);
                    const float e1 = gelu_erf_mufu(__uint_as_float(v[c8 * 8 + 2 * t + 1]) + bv[2 * t + 1]);
                    split_pair(e0, e1, hi[t], lo[t]);
                }
                const int chunk = (half * 4 + c8) ^ (r & 7);
                *reinterpret_cast<uint4*>(p_hi + chunk * 16) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                *reinterpret_cast<uint4*>(p_lo + chunk * 16) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_p(j));
        }
        mbar_wait(bar_o, 0);               // all MMAs of this CTA retired: ring, P and (for reads) the O accumulator are ours
        tc_fence_after();
    }

    if constexpr (!kCluster) {
        // ===== split variant: partial O (all 256 columns) -> global scratch, coalesced through the per-warp staging =====
        if (warp >= 2) {
            const int q = warp & 3, half = (warp - 2) >> 2;
            const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + 256;
            float* stg = reinterpret_cast<float*>(smem_gen + OFF_X) + (warp - 2) * 32 * EPI_PITCH;
            const int sub = lane >> 2, cq = (lane & 3) * 4;
            const size_t mpad = (size_t)((p.M + BM - 1) / BM) * BM;
            float* dst = p.partial + ((size_t)crank * mpad + m0 + q * 32) * C + half * 128;
#pragma unroll 1
            for (int c = 0; c < 128; c += EPI_CHUNK) {
                uint32_t r16[16];
                __syncwarp();
                tmem_ld16(trow + half * 128 + c, r16);
#pragma unroll
                for (int j = 0; j < 16; j += 4)
                    *reinterpret_cast<uint4*>(stg + lane * EPI_PITCH + j) = make_uint4(r16[j], r16[j + 1], r16[j + 2], r16[j + 3]);
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int rl = i * 8 + sub;
                    *reinterpret_cast<float4*>(dst + (size_t)rl * C + c + cq) = *reinterpret_cast<const float4*>(stg + rl * EPI_PITCH + cq);
                }
            }
        }
        if (p.cnt) {
            // ===== in-kernel reduce: x <- ((sum_s partial_s + b2) * gamma + x) * mask (+ post-ops) for this CTA's share of the tile's rows,
            //       partials summed in slice order with the expressions of mlp_reduce_post_kernel (bit-identical to the two-launch form)
            int* cnt = p.cnt + 2 * (m0 / BM);
            __threadfence();
            __syncthreads();
            if (threadIdx.x == 0) {
                __threadfence();           // cumulative: the CTA's partial stores (ordered before the barrier) precede the arrival
                atomicAdd(cnt, 1);
                int seen;
                do { asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(seen) : "l"(cnt) : "memory"); } while (seen < CSt);
                if (atomicAdd(cnt + 1, 1) == CSt - 1) { cnt[0] = 0; cnt[1] = 0; }      // every CTA of the tile has seen the full count
                __threadfence();
            }
            __syncthreads();
            if (warp >= 2) {
                constexpr int RPC = BM / CSt, RPW = RPC / 8;           // rows per CTA / per epilogue warp
                static_assert(RPW >= 1, "eight epilogue warps share a CTA's rows");
                const size_t slice = (size_t)((p.M + BM - 1) / BM) * BM * C;
                const int c0 = lane * 8, row0 = m0 + crank * RPC + (warp - 2) * RPW;
                float y[RPW][8];
#pragma unroll
                for (int i = 0; i < RPW; ++i) {
                    const size_t o = (size_t)(row0 + i) * C + c0;
                    const float4 a0 = __ldcg(reinterpret_cast<const float4*>(p.partial + o)), a1 = __ldcg(reinterpret_cast<const float4*>(p.partial + o + 4));
                    y[i][0] = a0.x; y[i][1] = a0.y; y[i][2] = a0.z; y[i][3] = a0.w; y[i][4] = a1.x; y[i][5] = a1.y; y[i][6] = a1.z; y[i][7] = a1.w;
                }
#pragma unroll 4
                for (int sl = 1; sl < CSt; ++sl) {
#pragma unroll
                    for (int i = 0; i < RPW; ++i) {
                        const size_t o = sl * slice + (size_t)(row0 + i) * C + c0;
                        const float4 v0 = __ldcg(reinterpret_cast<const float4*>(p.partial + o)), v1 = __ldcg(reinterpret_cast<const float4*>(p.partial + o + 4));
                        y[i][0] += v0.x; y[i][1] += v0.y; y[i][2] += v0.z; y[i][3] += v0.w; y[i][4] += v1.x; y[i][5] += v1.y; y[i][6] += v1.z; y[i][7] += v1.w;
                    }
                }
                float bb[8], gg[8];
                *reinterpret_cast<float4*>(bb) = __ldg(reinterpret_cast<const float4*>(p.b2 + c0)); *reinterpret_cast<float4*>(bb + 4) = __ldg(reinterpret_cast<const float4*>(p.b2 + c0 + 4));
                *reinterpret_cast<float4*>(gg) = __ldg(reinterpret_cast<const float4*>(p.gamma + c0)); *reinterpret_cast<float4*>(gg + 4) = __ldg(reinterpret_cast<const float4*>(p.gamma + c0 + 4));
#pragma unroll
                for (int i = 0; i < RPW; ++i) {
                    const int row = row0 + i;
                    if (row >= p.M) continue;
                    const size_t o = (size_t)row * C + c0;
                    const float mk = p.mask ? __ldg(p.mask + row) : 1.f;
                    float rr[8];
                    *reinterpret_cast<float4*>(rr) = *reinterpret_cast<const float4*>(p.x + o); *reinterpret_cast<float4*>(rr + 4) = *reinterpret_cast<const float4*>(p.x + o + 4);
#pragma unroll
                    for (int j = 0; j < 8; ++j) y[i][j] = ((y[i][j] + bb[j]) * gg[j] + rr[j]) * mk;
                    if (p.add_vec) {
                        float tt[8];
                        *reinterpret_cast<float4*>(tt) = __ldg(reinterpret_cast<const float4*>(p.add_vec + c0)); *reinterpret_cast<float4*>(tt + 4) = __ldg(reinterpret_cast<const float4*>(p.add_vec + c0 + 4));
#pragma unroll
                        for (int j = 0; j < 8; ++j) y[i][j] = (y[i][j] + tt[j]) * mk;
                    }
                    *reinterpret_cast<float4*>(p.x + o) = make_float4(y[i][0], y[i][1], y[i][2], y[i][3]);
                    *reinterpret_cast<float4*>(p.x + o + 4) = make_float4(y[i][4], y[i][5], y[i][6], y[i][7]);
                    if (!p.out_hi) continue;
                    if (p.post_ln_g) {
                        float sm = 0.f;
#pragma unroll
                        for (int j = 0; j < 8; ++j) sm += y[i][j];
                        const float mean = warp_sum<float>(sm) / (float)C;
                        float vv = 0.f;
#pragma unroll
                        for (int j = 0; j < 8; ++j) { y[i][j] -= mean; vv += y[i][j] * y[i][j]; }
                        const float inv = 1.0f / sqrtf(warp_sum<float>(vv) / (float)C + 1e-6f);
                        float g8[8], h8[8];
                        *reinterpret_cast<float4*>(g8) = __ldg(reinterpret_cast<const float4*>(p.post_ln_g + c0)); *reinterpret_cast<float4*>(g8 + 4) = __ldg(reinterpret_cast<const float4*>(p.post_ln_g + c0 + 4));
                        *reinterpret_cast<float4*>(h8) = __ldg(reinterpret_cast<const float4*>(p.post_ln_b + c0)); *reinterpret_cast<float4*>(h8 + 4) = __ldg(reinterpret_cast<const float4*>(p.post_ln_b + c0 + 4));
#pragma unroll
                        for (int j = 0; j < 8; ++j) y[i][j] = y[i][j] * inv * g8[j] + h8[j];
                    }
                    uint32_t hi[4], lo[4];
#pragma unroll
                    for (int t = 0; t < 4; ++t) split_pair(y[i][2 * t], y[i][2 * t + 1], hi[t], lo[t]);
                    *reinterpret_cast<uint4*>(p.out_hi + o) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                    *reinterpret_cast<uint4*>(p.out_lo + o) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                }
            }
        }
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        if (warp == 1) tmem_dealloc(tmem_base, 512);
        return;
    }
    // ===== cross-CTA reduction of the partial outputs (all threads take part in the cluster barriers) =====
    cluster_sync_all();                    // every CTA's weight ring is dead: it becomes the receive area
    if (warp >= 2) {
        const int q = warp & 3, half = (warp - 2) >> 2, r = q * 32 + lane;
        const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + 256;
        for (int d = 1; d < CS; ++d) {
            const int peer = (crank + d) % CS;                       // stagger the targets
            const int slot = crank < peer ? crank : crank - 1;       // my slot among the peer's three senders
            uint32_t v[32];
            __syncwarp();
            tmem_ld32(trow + peer * 64 + half * 32, v);
            const uint32_t local = smem_base + OFF_RING + slot * RECV_BYTES + r * 256;
            const uint32_t remote = mapa_u32(local, (uint32_t)peer);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int chunk = (half * 8 + i) ^ (r & 15);
                st_cluster_v4(remote + chunk * 16, __uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]),
                              __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
            }
        }
    }
    cluster_sync_all();                    // partials have landed (release / acquire at cluster scope)
    if (warp >= 2) {
        const int q = warp & 3, half = (warp - 2) >> 2, r = q * 32 + lane;
        const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + 256;
        float* stg = reinterpret_cast<float*>(smem_gen + OFF_X) + (warp - 2) * 32 * EPI_PITCH;
        const int sub = lane >> 2, cq = (lane & 3) * 4;
        float acc[32];
        {
            uint32_t v[32];
            __syncwarp();
            tmem_ld32(trow + crank * 64 + half * 32, v);
            // sum the four partials in source-rank order (own one at position crank): the same order in every CTA
#pragma unroll
            for (int t = 0; t < 32; ++t) acc[t] = 0.f;
            for (int src = 0; src < CS; ++src) {
                if (src == crank) {
#pragma unroll
                    for (int t = 0; t < 32; ++t) acc[t] += __uint_as_float(v[t]);
                } else {
                    const int slot = src < crank ? src : src - 1;
                    const uint8_t* row = smem_gen + OFF_RING + slot * RECV_BYTES + r * 256;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int chunk = (half * 8 + i) ^ (r & 15);
                        const float4 w = *reinterpret_cast<const float4*>(row + chunk * 16);
                        acc[4 * i] += w.x; acc[4 * i + 1] += w.y; acc[4 * i + 2] += w.z; acc[4 * i + 3] += w.w;
                    }
                }
            }
        }
        // epilogue 2 through per-warp staging: lanes along the row, coalesced residual reads and stores
        const int mrow0 = m0 + q * 32, col0 = crank * 64 + half * 32;
#pragma unroll
        for (int c = 0; c < 32; c += EPI_CHUNK) {
            __syncwarp();
#pragma unroll
            for (int j = 0; j < EPI_CHUNK; j += 4)
                *reinterpret_cast<float4*>(stg + lane * EPI_PITCH + j) = make_float4(acc[c + j], acc[c + j + 1], acc[c + j + 2], acc[c + j + 3]);
            __syncwarp();
            const int col = col0 + c + cq;
            const float4 b2 = __ldg(reinterpret_cast<const float4*>(p.b2 + col));
            const float4 g = __ldg(reinterpret_cast<const float4*>(p.gamma + col));
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int rl = i * 8 + sub, row = mrow0 + rl;
                if (row >= p.M) continue;
                float4 v = *reinterpret_cast<const float4*>(stg + rl * EPI_PITCH + cq);
                float* xp = p.x + (size_t)row * C + col;
                const float4 res = *reinterpret_cast<const float4*>(xp);
                const float mk = p.mask ? __ldg(p.mask + row) : 1.f;
                v.x = ((v.x + b2.x) * g.x + res.x) * mk; v.y = ((v.y + b2.y) * g.y + res.y) * mk;
                v.z = ((v.z + b2.z) * g.z + res.z) * mk; v.w = ((v.w + b2.w) * g.w + res.w) * mk;
                *reinterpret_cast<float4*>(xp) = v;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

__global__ void __cluster_dims__(CS, 1, 1) __launch_bounds__(NUM_THREADS, 1)
convnext_mlp_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                    const __grid_constant__ CUtensorMap map_w1_hi, const __grid_constant__ CUtensorMap map_w1_lo,
                    const __grid_constant__ CUtensorMap map_w2_hi, const __grid_constant__ CUtensorMap map_w2_lo,
                    const Params p) {
    convnext_mlp_body<true>(map_a_hi, map_a_lo, map_w1_hi, map_w1_lo, map_w2_hi, map_w2_lo, p);
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
convnext_mlp_thin_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                         const __grid_constant__ CUtensorMap map_w1_hi, const __grid_constant__ CUtensorMap map_w1_lo,
                         const __grid_constant__ CUtensorMap map_w2_hi, const __grid_constant__ CUtensorMap map_w2_lo,
                         const Params p) {
    convnext_mlp_body<false, HC_THIN>(map_a_hi, map_a_lo, map_w1_hi, map_w1_lo, map_w2_hi, map_w2_lo, p);
}

// map_w1_*: boxes of 64 weight rows here (128 in the other forms)
__global__ void __launch_bounds__(NUM_THREADS, 1)
convnext_mlp_thin64_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                           const __grid_constant__ CUtensorMap map_w1_hi, const __grid_constant__ CUtensorMap map_w1_lo,
                           const __grid_constant__ CUtensorMap map_w2_hi, const __grid_constant__ CUtensorMap map_w2_lo,
                           const Params p) {
    convnext_mlp_body<false, HC_THIN64>(map_a_hi, map_a_lo, map_w1_hi, map_w1_lo, map_w2_hi, map_w2_lo, p);
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
convnext_mlp_split_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                          const __grid_constant__ CUtensorMap map_w1_hi, const __grid_constant__ CUtensorMap map_w1_lo,
                          const __grid_constant__ CUtensorMap map_w2_hi, const __grid_constant__ CUtensorMap map_w2_lo,
                          const Params p) {
    convnext_mlp_body<false>(map_a_hi, map_a_lo, map_w1_hi, map_w1_lo, map_w2_hi, map_w2_lo, p);
}

// map_w1_* / map_w2_*: boxes of 256 weight rows
__global__ void __launch_bounds__(NUM_THREADS, 1)
convnext_mlp_split_wide_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                               const __grid_constant__ CUtensorMap map_w1_hi, const __grid_constant__ CUtensorMap map_w1_lo,
                               const __grid_constant__ CUtensorMap map_w2_hi, const __grid_constant__ CUtensorMap map_w2_lo,
                               const Params p) {
    convnext_mlp_body<false, HC, true>(map_a_hi, map_a_lo, map_w1_hi, map_w1_lo, map_w2_hi, map_w2_lo, p);
}

// ===== "TS" form of the split kernel: P = GELU(S) never touches shared memory =================================================
// The split kernel above is serial per CTA — phase 1 (all of S) -> GELU -> phase 2 — because P overwrites the a-tile: its tensor
// pipe is busy 43 % of the kernel (ncu, profiles/r1w_mlp_ncu_full_summary.txt). Here the epilogue warps write P back INTO the
// TMEM columns S came from (fp32 S block of 64 columns -> 32 columns of packed bf16 hi + 32 of lo) and phase 2 reads its A
// operand from TMEM (tcgen05.mma with [a_tmem], the form flash-attention kernels use for P.V). Nothing aliases the a-tile, so
//   * phase 1 runs hidden-half-major (nh outer): S[:, 0:128] is complete after 4 of its 8 weight units and its GELU overlaps
//     the MMAs of S[:, 128:256];
//   * phase 2 is K-block-major: its first half consumes P blocks 0,1 while the epilogue warps still produce blocks 2,3.
// TMEM A layout (M = 128, K-major bf16): lane = row, 32-bit column c of a K=16 slice holds elements (2c, 2c+1), 8 columns per
// slice. Per 16 hidden units g of block j: hi at column 64 j + 16 g, lo at + 8 — each warp only overwrites columns it has
// already read, so any number of warps per lane quarter can share a block.
STC_DEVINL void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
STC_DEVINL void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
          "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
STC_DEVINL void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

STC_DEVINL void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(map), "r"(src), "r"(c0), "r"(c1) : "memory");
}
STC_DEVINL void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> STC_DEVINL void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }

// Weight units u = 0..15 (32 KB each: 128 weight rows x 64 K, hi + lo). Phase 1 (u < 8) streams W1 through the 3-slot ring;
// phase 2 streams W2 through the ring AND through the dead a-tile (4 more slots, units 11..14), so that 7 of its 8 units are in
// flight as soon as S is complete (a unit takes ~2000 cycles to land when three share the SM's ~52 B/clk ingest).
struct UnitSlot { int ring; int idx; uint32_t par; };      // ring slot (idx, use parity) or a-region slot (idx, parity 0)
STC_DEVINL UnitSlot unit_slot(int u) {
    if (u >= 11 && u <= 14) return UnitSlot{0, u - 11, 0u};
    if (u == 15) return UnitSlot{1, 2, 1u};                 // fourth use of ring slot 2 (after u = 2, 5, 8)
    return UnitSlot{1, u % SLOTS, (uint32_t)((u / SLOTS) & 1)};
}

template <int NEPI>             // epilogue warps: 8 or 16 (NEPI / 4 per TMEM lane quarter)
__global__ void __launch_bounds__(64 + 32 * NEPI, 1)
convnext_mlp_ts_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                       const __grid_constant__ CUtensorMap map_w1_hi, const __grid_constant__ CUtensorMap map_w1_lo,
                       const __grid_constant__ CUtensorMap map_w2_hi, const __grid_constant__ CUtensorMap map_w2_lo,
                       const __grid_constant__ CUtensorMap map_part, const Params p) {
    using namespace tc;
    static_assert(NEPI == 8 || NEPI == 16, "two or four epilogue warps per TMEM lane quarter");
    constexpr int NP = NEPI / 4;              // warps sharing a lane quarter
    constexpr int CW = BK / NP;               // S columns of a 64-column block per warp (32 or 16)
    constexpr int OW = C / NP;                // output columns per warp in the final epilogue
    constexpr int NU = 2 * UNITS_PER_PHASE;
    if (p.trace && blockIdx.x == 0 && threadIdx.x == 64) p.trace[3] = clock64();
    pdl_trigger();
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t bar = smem_base + OFF_BAR;
    const uint32_t bar_a = bar, bar_o = bar + 16;
    auto bar_s = [&](int h) { return bar + 8 + 104u * h; };            // +8, +112
    auto full_bar = [&](int s) { return bar + 24 + 8u * s; };
    auto empty_bar = [&](int s) { return bar + 48 + 8u * s; };
    auto bar_p = [&](int j) { return bar + 72 + 8u * j; };
    auto afull_bar = [&](int i) { return bar + 128 + 8u * i; };
    const uint32_t tmem_slot = bar + 104;
    volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + OFF_BAR + 104);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int crank = (int)(blockIdx.x % CS);
    const int m0 = (int)(blockIdx.x / CS) * BM;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a_hi); tma_prefetch_desc(&map_a_lo); tma_prefetch_desc(&map_w1_hi);
        tma_prefetch_desc(&map_w1_lo); tma_prefetch_desc(&map_w2_hi); tma_prefetch_desc(&map_w2_lo); tma_prefetch_desc(&map_part);
        mbar_init(bar_a, 1); mbar_init(bar_s(0), 1); mbar_init(bar_s(1), 1); mbar_init(bar_o, 1);
        for (int s = 0; s < SLOTS; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int i = 0; i < 4; ++i) mbar_init(afull_bar(i), 1);
        for (int j = 0; j < HC / BK; ++j) mbar_init(bar_p(j), NEPI);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        STC_TRACE(4);
    }
    if (warp == 1) { tmem_alloc(tmem_slot, 512); STC_TRACE(5); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_gen;
    if (warp == 2) STC_TRACE(0);
    pdl_wait();
    if (warp == 2) STC_TRACE(1);

    if (warp == 0) {
        if (elect_one()) {
            mbar_expect_tx(bar_a, X_BYTES);
            for (int kb = 0; kb < C / BK; ++kb) {
                tma_load_2d(smem_base + OFF_X + kb * KBLK, &map_a_hi, bar_a, kb * BK, m0);
                tma_load_2d(smem_base + OFF_X + (C / BK + kb) * KBLK, &map_a_lo, bar_a, kb * BK, m0);
            }
            for (int u = 0; u < NU; ++u) {
                const UnitSlot us = unit_slot(u);
                uint32_t dst, fb;
                if (us.ring) {
                    mbar_wait(empty_bar(us.idx), us.par ^ 1);
                    dst = smem_base + OFF_RING + us.idx * UNIT; fb = full_bar(us.idx);
                } else {
                    if (us.idx == 0) mbar_wait(bar_s(1), 0);          // every phase-1 MMA has retired: the a-tile is dead
                    dst = smem_base + OFF_X + us.idx * UNIT; fb = afull_bar(us.idx);
                }
                if (p.trace && blockIdx.x == 0) p.trace[40 + u] = clock64();
                mbar_expect_tx(fb, UNIT);
                const int v = u % UNITS_PER_PHASE;
                if (u < UNITS_PER_PHASE) {          // W1[hidden rows, C]: hidden half nh = v / 4 (outer), K block kb = v % 4 of C
                    const int nh = v >> 2, kb = v & 3;
                    tma_load_2d(dst, &map_w1_hi, fb, kb * BK, crank * HC + nh * 128);
                    tma_load_2d(dst + KBLK, &map_w1_lo, fb, kb * BK, crank * HC + nh * 128);
                } else {                            // W2[C rows, hidden]: K block kb = v / 2 (outer) of this CTA's hidden slice, rows nh*128
                    const int kb = v >> 1, nh = v & 1;
                    tma_load_2d(dst, &map_w2_hi, fb, crank * HC + kb * BK, nh * 128);
                    tma_load_2d(dst + KBLK, &map_w2_lo, fb, crank * HC + kb * BK, nh * 128);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        constexpr uint32_t idesc = make_idesc_bf16(BM, 128);
        mbar_wait(bar_a, 0);
        STC_TRACE(2);
        for (int u = 0; u < NU; ++u) {
            const UnitSlot us = unit_slot(u);
            const int v = u % UNITS_PER_PHASE;
            const bool second = u >= UNITS_PER_PHASE;
            const int kb = second ? (v >> 1) : (v & 3), nh = second ? (v & 1) : (v >> 2);
            if (second && nh == 0) mbar_wait(bar_p(kb), 0);             // P block kb is in TMEM
            mbar_wait(us.ring ? full_bar(us.idx) : afull_bar(us.idx), us.par);
            tc_fence_after();
            STC_TRACE(8 + u);
            if (elect_one()) {
                const uint32_t st = us.ring ? smem_base + OFF_RING + us.idx * UNIT : smem_base + OFF_X + us.idx * UNIT;
                const uint64_t w_hi = make_smem_desc(st), w_lo = make_smem_desc(st + KBLK);
                if (!second) {
                    const uint32_t xk = smem_base + OFF_X + kb * KBLK;
                    const uint64_t a_hi = make_smem_desc(xk), a_lo = make_smem_desc(xk + (C / BK) * KBLK);
                    const uint32_t d = tmem_base + nh * 128;
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        const uint64_t adv = (uint64_t)((k * UMMA_K * 2) >> 4);
                        umma_bf16(d, a_lo + adv, w_hi + adv, idesc, (kb | k) != 0);
                        umma_bf16(d, a_hi + adv, w_lo + adv, idesc, 1);
                        umma_bf16(d, a_hi + adv, w_hi + adv, idesc, 1);
                    }
                } else {
                    const uint32_t d = tmem_base + 256 + nh * 128;
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        const uint64_t adv = (uint64_t)((k * UMMA_K * 2) >> 4);
                        const uint32_t p_hi = tmem_base + kb * BK + k * UMMA_K, p_lo = p_hi + UMMA_K / 2;
                        umma_bf16_ts(d, p_lo, w_hi + adv, idesc, (kb | k) != 0);
                        umma_bf16_ts(d, p_hi, w_lo + adv, idesc, 1);
                        umma_bf16_ts(d, p_hi, w_hi + adv, idesc, 1);
                    }
                }
                if (us.ring) umma_commit(empty_bar(us.idx));
                if (!second && kb == C / BK - 1) umma_commit(bar_s(nh));          // S[:, 128 nh .. +128) complete
                if (u == NU - 1) umma_commit(bar_o);                               // partial O complete; every smem operand is dead
            }
            __syncwarp();
        }
    } else {
        // ===== epilogue 1: P = split(GELU(S + b1)), in place in TMEM =====
        const int q = warp & 3, part = (warp - 2) >> 2;
        const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
        for (int j = 0; j < HC / BK; ++j) {
            if ((j & 1) == 0) { mbar_wait(bar_s(j >> 1), 0); tc_fence_after(); if (warp == 2) STC_TRACE(24 + (j >> 1)); }
#pragma unroll
            for (int g = 0; g < CW / 16; ++g) {
                const int col = j * BK + part * CW + g * 16;
                uint32_t v[16], o[16];
                __syncwarp();
                tmem_ld16(trow + col, v);
                const float* b1 = p.b1 + crank * HC + col;
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    const float e0 = gelu_erf_mufu(__uint_as_float(v[2 * t]) + __ldg(b1 + 2 * t));
                    const float e1 = gelu_erf_mufu(__uint_as_float(v[2 * t + 1]) + __ldg(b1 + 2 * t + 1));
                    split_pair(e0, e1, o[t], o[8 + t]);
                }
                tmem_st16(trow + col, o);
            }
            tmem_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_p(j));
            if (warp == 2) STC_TRACE(26 + j);
        }
        mbar_wait(bar_o, 0);               // all MMAs of this CTA retired: the a-tile region (staging below) and the O accumulator are ours
        tc_fence_after();
        if (warp == 2) STC_TRACE(30);
        // ===== partial O (256 columns) -> global scratch by TMA: 32 rows x 32 columns (one 128-byte swizzled row per lane) per
        //       store, two staging buffers per warp so that the copy of one box overlaps the TMEM read of the next =====
        const uint32_t orow = trow + 256 + part * OW;
        const uint32_t stg = smem_base + OFF_X + (uint32_t)(warp - 2) * 8192u;
        const int mpad = ((p.M + BM - 1) / BM) * BM;
        const int grow = crank * mpad + m0 + q * 32;
#pragma unroll 1
        for (int c = 0; c < OW; c += 32) {
            const uint32_t buf = stg + (uint32_t)((c >> 5) & 1) * 4096u;
            if (c >= 64) { if (lane == 0) bulk_wait_read<1>(); __syncwarp(); }        // the store that last read this buffer has drained
            uint32_t v[32];
            tmem_ld32(orow + c, v);
#pragma unroll
            for (int ch = 0; ch < 8; ++ch)
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(buf + (uint32_t)lane * 128u + (uint32_t)((ch ^ (lane & 7)) * 16)),
                             "r"(v[4 * ch]), "r"(v[4 * ch + 1]), "r"(v[4 * ch + 2]), "r"(v[4 * ch + 3]) : "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) { tma_store_2d(&map_part, buf, part * OW + c, grow); bulk_commit(); }
        }
        if (lane == 0) bulk_wait_read<0>();
        __syncwarp();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 2) STC_TRACE(31);
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// x <- ((p0 + p1 + p2 + p3 + b2) * gamma + x) * mask : the four hidden-slice partials in rank order (deterministic).
// One warp per row (lane owns 8 consecutive channels), so what FOLLOWS the block in the graphs can ride along instead of
// costing a launch of its own:
//   add_vec : x <- (x + add_vec) * mask           (time conditioning after a ConvNeXt block, cpp-side: vector_estimator graph)
//   out     : the next layer's tensor-core operand — LayerNorm(x) (pre-LN of an attention layer; ln_g / ln_b) or x itself
//             (ln_g == null; input of an output projection) as split-bf16
// WIDE (the thin forms: 8 or 16 hidden slices, small batches): the partial loads of up to eight slices are in flight together —
// those launches wait for L2 round trips, not for bandwidth (batch-1 latency 5.22 -> 4.96 ms); with four slices and 4 736 rows the
// extra registers cost occupancy instead (10.40 -> 10.86 ms/step), so the two-deep loop stays there.
template <bool WIDE>
__global__ void __launch_bounds__(256)
mlp_reduce_post_kernel(const float* __restrict__ partial, size_t slice, const float* __restrict__ b2, const float* __restrict__ gamma,
                  const float* __restrict__ mask, float* __restrict__ x, int M, const float* __restrict__ add_vec,
                  const float* __restrict__ ln_g, const float* __restrict__ ln_b, float eps,
                  __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo, int nslice) {
    pdl_trigger(); pdl_wait();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + warp;
    if (row >= M) return;
    const int c0 = lane * 8;
    const size_t i = (size_t)row * C + c0;
    float y[8];
    {
        const float4 a0 = *reinterpret_cast<const float4*>(partial + i), a1 = *reinterpret_cast<const float4*>(partial + i + 4);
        y[0] = a0.x; y[1] = a0.y; y[2] = a0.z; y[3] = a0.w; y[4] = a1.x; y[5] = a1.y; y[6] = a1.z; y[7] = a1.w;
    }
    if constexpr (!WIDE) {
#pragma unroll 4
        for (int s = 1; s < nslice; ++s) {
            const float4 v0 = *reinterpret_cast<const float4*>(partial + s * slice + i), v1 = *reinterpret_cast<const float4*>(partial + s * slice + i + 4);
            y[0] += v0.x; y[1] += v0.y; y[2] += v0.z; y[3] += v0.w; y[4] += v1.x; y[5] += v1.y; y[6] += v1.z; y[7] += v1.w;
        }
    } else
    for (int s0 = 1; s0 < nslice; s0 += 8) {           // eight slices' loads in flight at a time, summed in slice order
        float4 q0[8], q1[8];
#pragma unroll
        for (int s = 0; s < 8; ++s)
            if (s0 + s < nslice) {
                q0[s] = *reinterpret_cast<const float4*>(partial + (s0 + s) * slice + i);
                q1[s] = *reinterpret_cast<const float4*>(partial + (s0 + s) * slice + i + 4);
            }
#pragma unroll
        for (int s = 0; s < 8; ++s)
            if (s0 + s < nslice) {
                y[0] += q0[s].x; y[1] += q0[s].y; y[2] += q0[s].z; y[3] += q0[s].w; y[4] += q1[s].x; y[5] += q1[s].y; y[6] += q1[s].z; y[7] += q1[s].w;
            }
    }
    const float mk = mask ? __ldg(mask + row) : 1.f;
    float bb[8], gg[8], rr[8];
    *reinterpret_cast<float4*>(bb) = __ldg(reinterpret_cast<const float4*>(b2 + c0)); *reinterpret_cast<float4*>(bb + 4) = __ldg(reinterpret_cast<const float4*>(b2 + c0 + 4));
    *reinterpret_cast<float4*>(gg) = __ldg(reinterpret_cast<const float4*>(gamma + c0)); *reinterpret_cast<float4*>(gg + 4) = __ldg(reinterpret_cast<const float4*>(gamma + c0 + 4));
    *reinterpret_cast<float4*>(rr) = *reinterpret_cast<const float4*>(x + i); *reinterpret_cast<float4*>(rr + 4) = *reinterpret_cast<const float4*>(x + i + 4);
#pragma unroll
    for (int j = 0; j < 8; ++j) y[j] = ((y[j] + bb[j]) * gg[j] + rr[j]) * mk;
    if (add_vec) {
        float tt[8];
        *reinterpret_cast<float4*>(tt) = __ldg(reinterpret_cast<const float4*>(add_vec + c0)); *reinterpret_cast<float4*>(tt + 4) = __ldg(reinterpret_cast<const float4*>(add_vec + c0 + 4));
#pragma unroll
        for (int j = 0; j < 8; ++j) y[j] = (y[j] + tt[j]) * mk;
    }
    *reinterpret_cast<float4*>(x + i) = make_float4(y[0], y[1], y[2], y[3]);
    *reinterpret_cast<float4*>(x + i + 4) = make_float4(y[4], y[5], y[6], y[7]);
    if (!out_hi) return;
    if (ln_g) {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) s += y[j];
        const float mean = warp_sum<float>(s) / (float)C;
        float v = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) { y[j] -= mean; v += y[j] * y[j]; }
        const float inv = 1.0f / sqrtf(warp_sum<float>(v) / (float)C + eps);
        float g8[8], h8[8];
        *reinterpret_cast<float4*>(g8) = __ldg(reinterpret_cast<const float4*>(ln_g + c0)); *reinterpret_cast<float4*>(g8 + 4) = __ldg(reinterpret_cast<const float4*>(ln_g + c0 + 4));
        *reinterpret_cast<float4*>(h8) = __ldg(reinterpret_cast<const float4*>(ln_b + c0)); *reinterpret_cast<float4*>(h8 + 4) = __ldg(reinterpret_cast<const float4*>(ln_b + c0 + 4));
#pragma unroll
        for (int j = 0; j < 8; ++j) y[j] = y[j] * inv * g8[j] + h8[j];
    }
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) tc::split_pair(y[2 * t], y[2 * t + 1], hi[t], lo[t]);
    *reinterpret_cast<uint4*>(out_hi + i) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(out_lo + i) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

// The plain reduce (no post-ops): one float4 per thread — twice the threads of the row-wise kernel above and ~3 us faster
// (4.2 vs 7.0 us at 4 736 rows, profiles/r1w_mlp_ncu_full_summary.txt), so blocks with nothing folded in keep this form.
template <bool WIDE>
__global__ void __launch_bounds__(256)
mlp_reduce_kernel(const float* __restrict__ partial, size_t slice, const float* __restrict__ b2, const float* __restrict__ gamma,
                  const float* __restrict__ mask, float* __restrict__ x, int M, int nslice) {
    pdl_trigger(); pdl_wait();
    const size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i >= (size_t)M * C) return;
    const int row = (int)(i / C), col = (int)(i % C);
    float4 acc;
    if constexpr (WIDE) {
        // every partial's load is in flight before the first add (one L2 round trip for up to 16 hidden slices), summed in slice order
        float4 pv[16];
#pragma unroll
        for (int s = 0; s < 16; ++s) if (s < nslice) pv[s] = *reinterpret_cast<const float4*>(partial + s * slice + i);
        acc = pv[0];
#pragma unroll
        for (int s = 1; s < 16; ++s) if (s < nslice) { acc.x += pv[s].x; acc.y += pv[s].y; acc.z += pv[s].z; acc.w += pv[s].w; }
    } else {
        acc = *reinterpret_cast<const float4*>(partial + i);
#pragma unroll 4
        for (int s = 1; s < nslice; ++s) {
            const float4 v = *reinterpret_cast<const float4*>(partial + s * slice + i);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
    }
    const float4 b = __ldg(reinterpret_cast<const float4*>(b2 + col)), g = __ldg(reinterpret_cast<const float4*>(gamma + col));
    const float4 r = *reinterpret_cast<const float4*>(x + i);
    const float mk = mask ? __ldg(mask + row) : 1.f;
    acc.x = ((acc.x + b.x) * g.x + r.x) * mk; acc.y = ((acc.y + b.y) * g.y + r.y) * mk;
    acc.z = ((acc.z + b.z) * g.z + r.z) * mk; acc.w = ((acc.w + b.w) * g.w + r.w) * mk;
    *reinterpret_cast<float4*>(x + i) = acc;
}

}  // namespace mlp
}  // namespace stc
