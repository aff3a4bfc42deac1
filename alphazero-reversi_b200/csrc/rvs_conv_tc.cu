// rvs_conv_tc.cu -- 3x3 convolution C->C (+ folded BN bias, optional residual, ReLU) as an
// implicit GEMM on the 5th-generation tensor cores: TMA -> shared memory -> tcgen05.mma -> TMEM ->
// tcgen05.ld epilogue.  sm_100a only.
//
// GEMM view per CTA PAIR (cta_group::2):  D[256 pixels, C couts] += A_tap[256 pixels, cin] * W_tap[C couts, cin]^T
//   M = 2 x 128: each CTA supplies one "tile" = two boards, rows ordered (y, board, x): with this order a
//       vertical tap shift dy is a shift by 16 rows = 2048 B, i.e. a whole number of 1024-byte swizzle atoms.
//   N = C couts; each CTA holds HALF of the weight rows (resident for C <= 128, streamed for C = 256).
//   K = 9 taps x cin, consumed as 3 (dx) x cin/64 stages; each stage is ONE TMA box of the
//       activations: 10 rows of y (halo -1..8) x 2 boards x 8 x (shifted by dx) x 64 channels,
//       out-of-bounds rows/columns zero-filled by the TMA unit = the convolution's zero padding.
//       The three vertical taps of that dx reuse the same box at row offsets 0 / 16 / 32.
// Activations live in HBM in the same tiled order [tile][y][board][x][c] (rvs_conv_tc.cuh:
// act_row), so accumulator row r of a tile is simply row r of the output tile.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer (leader CTA only),
// warps 2-5 = epilogue (TMEM lane quarter = warp_id % 4).  Two accumulators in TMEM ping-pong so that the
// epilogue of tile i overlaps the MMAs of tile i+1.  (The first-generation 1-CTA kernel -- N = 64 per CTA,
// tensor pipe at 31 % of its throughput, profiles/superseded/ncu_conv_tc1_r1.txt -- was removed in round 2.)
#include "rvs_conv_tc.cuh"

#include <cuda.h>
#include <stdlib.h>

#include "rvs_common.cuh"

namespace rvs {

namespace {

constexpr int kTileRows = 128;              // pixels per tile (2 boards)
constexpr int kABytes = 10 * 2 * 8 * 128;   // one activation stage of the streamed (256-filter) kernel: 160 rows x 128 B
// Resident-weight kernels (C <= 128): ONE activation box per 64-channel chunk serves all nine taps.  The box is
// 10 rows of y (halo -1..8) x 2 boards x 16 x-slots (x = -1..14; x < 0 and x > 7 are zero-filled by the TMA unit)
// x 64 channels, so a (y, board) line is exactly two 1024-byte swizzle atoms and a horizontal tap shift dx is a
// start-address offset of dx rows (128 B) INSIDE the atom (make_desc_x).  Compared with one shifted box per dx this fetches every input element once instead of three
// times (L2 -> SM traffic per tile 40 KB instead of 120 KB per chunk pair) and writes 80 KB instead of 120 KB of
// shared memory per tile -- the 128-filter layer was bound by exactly these two.
#ifndef RVS_CONV_XSLOTS
#define RVS_CONV_XSLOTS 16
#endif
constexpr int kXSlots = RVS_CONV_XSLOTS;
constexpr int kABytesX = 10 * 2 * kXSlots * 128;  // 40 KB
constexpr int kThreads = 192;   // streamed 256-filter kernel: warp 0 producer, warp 1 MMA issuer, warps 2-5 epilogue

struct Impl {
    CUtensorMap w_map2;   // box of C/2 weight rows (one CTA's half of a tap)
    int cin = 0;
    // TMA descriptors of the activation buffers this layer has been launched on, keyed by (base pointer, capacity):
    // 3 ping-pong buffers x 2 half-batches of a pipelined search
    static constexpr int kActSlots = 8;
    const void* act_ptr[kActSlots] = {};
    int64_t act_cap[kActSlots] = {};
    CUtensorMap act_map[kActSlots];
    int n_act = 0;
};

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeFn get_encode() {
    static EncodeFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeFn)p;
    }
    return fn;
}

// ---- PTX wrappers ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}
// 256-bit global accesses (LDG/STG.E.ENL2.256 on sm_100a): one full 32-byte sector per instruction
__device__ __forceinline__ void ldg256(const void* p, uint32_t* v) {
    asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "l"(p));
}
// the same through L2 only: for data that the TMA unit (async proxy) wrote, which an earlier L1 line would not reflect
__device__ __forceinline__ void ldg256_cg(const void* p, uint32_t* v) {
    asm volatile("ld.global.cg.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "l"(p));
}
__device__ __forceinline__ void stg256(void* p, const uint32_t* v) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]),
                 "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// shared-memory matrix descriptor, K-major, SWIZZLE_128B (cute::UMMA::SmemDescriptor):
//   [0,14) start address >> 4 | [16,30) leading byte offset >> 4 (=1, unused for swizzled K-major)
//   [32,46) stride byte offset >> 4 (1024 B between 8-row groups) | [46,48) version = 1 (sm_100)
//   [61,64) layout type = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// the same with 2048 B between 8-row groups (16 x-slots per (y, board) line).  The start address may sit 1 or 2 rows
// (128 / 256 B) into a swizzle atom (horizontal tap shift): the tensor core applies the 128-byte swizzle to the
// ABSOLUTE shared-memory address bits, exactly like the TMA unit that wrote the box, so no descriptor field has to
// describe the phase.  Measured on B200: with the descriptor's base-offset field [49,52) set to the row phase the
// results are wrong, with 0 all network goldens pass (tests/test_gpu_net.py).
__device__ __forceinline__ uint64_t make_desc_x(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | (1ull << 16) | ((uint64_t)((kXSlots * 128) >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D=f32 [4,6)=1, A=bf16 [7,10)=1, B=bf16
// [10,13)=1, A/B K-major (bits 15,16 = 0), N>>3 at [17,23), M>>4 at [24,29)  -> Cfg2::IDESC / Cfg2S::IDESC

// =============================================================================================
// CTA pairs (cta_group::2).  Profiling a 1-CTA kernel showed the tensor pipe busy 73 % of
// the time at only 31 % of its throughput: with N = 64 every MMA streams 4 KB of A + 2 KB of B
// from shared memory for 32 cycles of math (192 B/cycle > the 128 B/cycle SMEM port).  Here a CTA
// pair computes D[256 px, C couts]: each CTA supplies its own 128-pixel tile (A) and HALF of the
// weight rows (B, still resident), so one MMA reads 4 KB + 2 KB per CTA for 64 cycles of math
// (96 B/cycle) and every activation byte is fetched once per cout instead of twice.
//   * both CTAs run a TMA producer (own A tile, own W half) that signals the LEADER's barriers
//     (cp.async.bulk.tensor ... .cta_group::2 with the peer bit of the barrier address cleared);
//   * only the leader (cluster rank 0) issues tcgen05.mma.cta_group::2 and commits with
//     .multicast::cluster so that "stage free" / "accumulator ready" arrive in both CTAs;
//   * both CTAs run the epilogue on their own TMEM (128 lanes x C columns, double buffered); the
//     peer's epilogue warps arrive remotely on the leader's "accumulator free" barrier.
// =============================================================================================
// programmatic dependent launch: the next layer's prologue (barrier init, TMEM alloc, 144 KB of
// weights) overlaps the tail of this layer; activations are only touched after pdl_wait()
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;  // cute::Sm100MmaPeerBitMask: address of the same offset in the even CTA
__device__ __forceinline__ void tma2_load_2d(const CUtensorMap* map, uint32_t leader_bar, uint32_t dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(map), "r"(leader_bar & kPeerBitMask), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma2_load_5d(const CUtensorMap* map, uint32_t leader_bar, uint32_t dst, int c0, int c1, int c2,
                                             int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst),
        "l"(map), "r"(leader_bar & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}
__device__ __forceinline__ void tc2_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}
__device__ __forceinline__ void tc2_commit_mc(uint32_t bar) {  // arrive on `bar` in BOTH CTAs of the pair
    asm volatile(
        "{\n\t.reg .b16 m;\n\tmov.b16 m, 3;\n\t"
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n\t}" ::"r"(bar)
        : "memory");
}
// Barrier traffic inside the pair carries NO generic-proxy memory dependency: operands travel TMA -> shared memory ->
// tensor core (async proxy, ordered by complete_tx / tcgen05.commit) and results tensor core -> TMEM -> registers
// (ordered by tcgen05.fence / tcgen05.wait::ld).  So the remote arrive and the waits use the default CTA-scope
// semantics, as CUTLASS's ClusterBarrier does.  The `.release.cluster` / `.acquire.cluster` forms used in round 1
// compiled to MEMBAR.ALL.GPU + ERRBAR per arrive and to CCTL.IVALL (invalidate all of L1) per successful wait: a
// GPU-wide fence per epilogue warp per tile that waited for the tile's 32 KB of output stores to drain (ncu:
// stall_membar 17 % of the epilogue warps' samples).
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar) {  // arrive on the leader CTA's barrier from either CTA
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, 0;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}" ::"r"(bar)
        : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}

// HEAD variant (last tower layer): the policy / value 1x1 convolutions + BN + ReLU
// (network.py:104-105, 111-112) are three dot products over the output row the epilogue already
// holds in registers, so the 67 MB activation write and its re-read by the head kernel disappear;
// weights travel in the kernel parameter space (constant cache, warp-uniform reads).
using HeadW = ConvHeadW;  // rvs_conv_tc.cuh

#ifdef RVS_CONV_PROBE  // debug variant only (tools/probe_conv.py): where the roles of the 128-filter kernel wait
__device__ long long g_conv_probe[148 * 16];
__device__ long long g_conv_layer[148 * 48];  // tower kernel's MMA issuer: clock at the start of every layer
#define PROBE_T0() const long long _pt = clock64()
#define PROBE_ADD(x) (x) += clock64() - _pt
#else
#define PROBE_T0()
#define PROBE_ADD(x)
#endif

template <int C, int CIN>
struct Cfg2 {
    static constexpr int KC = CIN / 64;
    static constexpr int NH = C / 2;                        // weight rows (couts) held by each CTA
    static constexpr int W_TILE = NH * 128;                 // bytes of one (tap, kc) weight tile per CTA
    static constexpr int W_TILES = 9 * KC;
    static constexpr int STAGES = (C == 128 && CIN == 128) ? (kXSlots <= 10 ? 3 : 2) : (C == 128 ? 3 : 4);  // 128->128: 144 KB weights + 2 x 40 KB = 224 KB
    static constexpr int TMEM_COLS = 2 * C;                 // two accumulators of C fp32 columns
    // K = 16 steps per 64-channel chunk.  The 64 -> 128 instantiation is the network's FIRST layer: only
    // channels 0..2 of its input tiles are non-zero, so the steps over channels 16..63 would multiply zeros
    static constexpr int KSTEPS = (C == 128 && CIN == 64) ? 1 : 4;
    static constexpr int EW = 4;                                // epilogue warps (4 or 8, see the epilogue's comment)
    static constexpr int THREADS = 64 + 32 * EW;                // warp 0 TMA producer, warp 1 MMA issuer, then the epilogue
    static constexpr int SMEM = W_TILES * W_TILE + STAGES * kABytesX + 1024 /*align*/ + 1280 /*barriers, bias*/;
    static constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(C >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
};

template <int C, int CIN, bool HEAD, bool FRESH>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(Cfg2<C, CIN>::THREADS, 1)
conv3x3_tc2_kernel(const __grid_constant__ CUtensorMap a_map, const __grid_constant__ CUtensorMap w_map,
                   const __nv_bfloat16* __restrict__ residual, __nv_bfloat16* __restrict__ out,
                   const float* __restrict__ bias, int n_tiles_arg, const __grid_constant__ HeadW head,
                   float* __restrict__ feat, const int* __restrict__ n_boards_dev, int rev) {
    using K = Cfg2<C, CIN>;
    // Compacted leaf batches: the number of boards is only known on the device.  Normally it was written at least two
    // kernels upstream and is read here at once (visible even when this launch overlaps its predecessor's tail).  FRESH:
    // the kernel just before this one wrote it (the first layer after the tree step of a wave-1 search, which releases
    // its dependents early), so every role re-reads it after its griddepcontrol.wait.  That is a template parameter and
    // not a flag because a trip count that comes from a volatile load is no longer a UNIFORM value for the compiler: the
    // MMA issuer's descriptor arithmetic then leaves the uniform datapath (5 R2UR per tcgen05.mma) and a tower layer ran
    // 4 % slower; the first layer issues 9 MMAs per tile and does not care.
    int n_tiles = (n_boards_dev && !FRESH) ? (*n_boards_dev + 1) >> 1 : n_tiles_arg;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* gen = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t w_s = base;
    const uint32_t a_s = base + K::W_TILES * K::W_TILE;
    unsigned char* tail = gen + K::W_TILES * K::W_TILE + K::STAGES * kABytesX;
    uint64_t* bars = reinterpret_cast<uint64_t*>(tail);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tail + 192);
    float* sbias = reinterpret_cast<float*>(tail + 256);  // [C] <= 256 floats
    const uint32_t bar0 = smem_u32(bars);
    auto FULL = [&](int s) { return bar0 + 8u * s; };                       // used in the leader
    auto EMPTY = [&](int s) { return bar0 + 8u * (K::STAGES + s); };        // both CTAs (multicast commit)
    const uint32_t WFULL = bar0 + 8u * (2 * K::STAGES);                     // leader
    auto ACC_FULL = [&](int a) { return bar0 + 8u * (2 * K::STAGES + 1 + a); };   // both CTAs
    auto ACC_EMPTY = [&](int a) { return bar0 + 8u * (2 * K::STAGES + 3 + a); };  // leader, one arrival per epilogue warp of both CTAs

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
    int n_iters = (n_tiles + 2 * n_pairs - 1) / (2 * n_pairs);  // same trip count in both CTAs of a pair
    auto wait_for_inputs = [&]() {
        pdl_wait();
        if (FRESH && n_boards_dev) {
            n_tiles = (__ldcg(n_boards_dev) + 1) >> 1;
            n_iters = (n_tiles + 2 * n_pairs - 1) / (2 * n_pairs);
        }
    };

    if (threadIdx.x == 0) {
        for (int s = 0; s < K::STAGES; ++s) { mbar_init(FULL(s), 1); mbar_init(EMPTY(s), 1); }
        mbar_init(WFULL, 1);
        for (int a = 0; a < 2; ++a) { mbar_init(ACC_FULL(a), 1); mbar_init(ACC_EMPTY(a), 2 * K::EW); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < C; i += K::THREADS) sbias[i] = bias[i];
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(K::TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();  // barriers of both CTAs are initialised before anyone signals across the pair
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) pdl_launch_dependents();

    if (warp == 0) {
        if (lane == 0) {  // ===== TMA producer (both CTAs) =====
            if (rank == 0) mbar_expect_tx(WFULL, 2 * K::W_TILES * K::W_TILE);  // both halves report to the leader
            for (int tap = 0; tap < 9; ++tap)
                for (int kc = 0; kc < K::KC; ++kc)
                    tma2_load_2d(&w_map, WFULL, w_s + (tap * K::KC + kc) * K::W_TILE, kc * 64, tap * C + (int)rank * K::NH);
            wait_for_inputs();  // weights are constants; activations come from the previous layer
            int stage = 0, phase = 0;
            long long p_empty = 0; (void)p_empty;
            for (int it = 0; it < n_iters; ++it) {
                const int tile = ((rev ? n_iters - 1 - it : it) * n_pairs + pair) * 2 + (int)rank;  // may be >= n_tiles: TMA zero-fills
                for (int kc = 0; kc < K::KC; ++kc) {
                    { PROBE_T0(); mbar_wait_cluster(EMPTY(stage), phase ^ 1); PROBE_ADD(p_empty); }
                    if (rank == 0) mbar_expect_tx(FULL(stage), 2 * kABytesX);
                    tma2_load_5d(&a_map, FULL(stage), a_s + stage * kABytesX, kc * 64, -1, 0, -1, tile);
                    if (++stage == K::STAGES) { stage = 0; phase ^= 1; }
                }
            }
#ifdef RVS_CONV_PROBE
            if (!HEAD && CIN == 128) g_conv_probe[blockIdx.x * 16 + 0] = p_empty;
#endif
        }
    } else if (warp == 1) {
        if (lane == 0 && rank == 0) {  // ===== MMA issuer (leader only) =====
            long long p_w = 0, p_acc = 0, p_full = 0, p_start = 0; (void)p_w; (void)p_acc; (void)p_full; (void)p_start;
#ifdef RVS_CONV_PROBE
            p_start = clock64();
#endif
            wait_for_inputs();
            { PROBE_T0(); mbar_wait_cluster(WFULL, 0); PROBE_ADD(p_w); }
            int stage = 0, phase = 0;
            for (int it = 0; it < n_iters; ++it) {
                const int acc = it & 1;
                { PROBE_T0(); mbar_wait_cluster(ACC_EMPTY(acc), ((it >> 1) & 1) ^ 1); PROBE_ADD(p_acc); }
                tc_fence_after();
                const uint32_t d = tmem_base + (uint32_t)(acc * C);
                uint32_t accum = 0;
                for (int kc = 0; kc < K::KC; ++kc) {
                    { PROBE_T0(); mbar_wait_cluster(FULL(stage), phase); PROBE_ADD(p_full); }
                    tc_fence_after();
                    const uint32_t a0 = a_s + stage * kABytesX;
#pragma unroll
                    for (int dy = 0; dy < 3; ++dy) {
#pragma unroll
                        for (int dx = 0; dx < 3; ++dx) {
                            const uint32_t wt = w_s + ((dy * 3 + dx) * K::KC + kc) * K::W_TILE;
                            // tap (dy, dx): input pixel (y + dy - 1, x + dx - 1) = box row ((y + dy) * 2 + board) * 16 + x + dx
                            const uint32_t at = a0 + (uint32_t)(dy * 2 * kXSlots + dx) * 128u;
#pragma unroll
                            for (int k = 0; k < K::KSTEPS; ++k) {
                                tc2_mma(d, make_desc_x(at + k * 32), make_desc(wt + k * 32), K::IDESC, accum);
                                accum = 1;
                            }
                        }
                    }
                    tc2_commit_mc(EMPTY(stage));
                    if (++stage == K::STAGES) { stage = 0; phase ^= 1; }
                }
                tc2_commit_mc(ACC_FULL(acc));
            }
#ifdef RVS_CONV_PROBE
            if (!HEAD && CIN == 128) {
            g_conv_probe[blockIdx.x * 16 + 1] = p_w;
            g_conv_probe[blockIdx.x * 16 + 2] = p_acc;
            g_conv_probe[blockIdx.x * 16 + 3] = p_full;
            g_conv_probe[blockIdx.x * 16 + 4] = clock64() - p_start;
            g_conv_probe[blockIdx.x * 16 + 7] = n_iters;
            }
#endif
        }
    } else {  // ===== epilogue (both CTAs, own tile) =====
        // Warp (q, hh) converts TMEM lane quarter q (32 pixel rows; hardware rule: warp id % 4) x channel slice hh.
        // FOUR warps (one slice each, EW = 4): a tile's conversion is ~900 instructions that one warp per scheduler runs at
        // 0.25 IPC, ~3000 cycles (~5000 with a residual) against ~5000 for the tile's 72 MMAs.  Eight warps (EW = 8, two
        // slices) finish it in half the time, but measured on B200 the MMAs themselves then take 81 instead of 70 cycles
        // each and a tower layer gets 3 % SLOWER (the SM is at its power limit under tensor load: what the other pipes
        // gain, the tensor pipe loses); the epilogue-paced first layer did not gain either (24.3 vs 23.7 us).
        constexpr int CH = C / (K::EW / 4);  // channels per epilogue warp
        const int q = warp & 3;
        const int hh = (warp - 2) >> 2;
        const int row = q * 32 + lane;
        long long p_af = 0, p_e0 = 0; (void)p_af; (void)p_e0;
        wait_for_inputs();
#ifdef RVS_CONV_PROBE
        p_e0 = clock64();
#endif
        for (int it = 0; it < n_iters; ++it) {
            const int acc = it & 1;
            const int tile = ((rev ? n_iters - 1 - it : it) * n_pairs + pair) * 2 + (int)rank;
            const size_t off = ((size_t)tile * kTileRows + row) * C;
            const bool live = tile < n_tiles;
            if constexpr (HEAD) {
                // The three head planes are dot products over ALL channels of a row in a fixed order: the hh = 0 warp of each
                // quarter walks the whole row, its hh = 1 partner only releases the accumulator.
                // plane = sum(channels 0..63) + sum(channels 64..127), each in ascending order (the whole-network kernel sums the
                // two halves in two warps: same order, same bits)
                float hd0 = 0.f, hd1 = 0.f, hd2 = 0.f, hs0 = 0.f, hs1 = 0.f, hs2 = 0.f;
                uint32_t res[C / 2];
                if (hh == 0) {
                    if (residual && live) {
#pragma unroll
                        for (int i = 0; i < C / 16; ++i) ldg256(residual + off + i * 16, res + i * 8);
                    } else {
#pragma unroll
                        for (int i = 0; i < C / 2; ++i) res[i] = 0u;
                    }
                }
                { PROBE_T0(); mbar_wait_cluster(ACC_FULL(acc), (it >> 1) & 1); PROBE_ADD(p_af); }
                tc_fence_after();
                if (hh == 0) {
#pragma unroll
                    for (int h = 0; h < C / 64; ++h) {
                        uint32_t v[64];
                        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * C + h * 64);
                        tc_ld32(taddr, v);
                        tc_ld32(taddr + 32, v + 32);
                        tc_wait_ld();
                        if (h == C / 64 - 1) {  // all TMEM reads of this accumulator are done
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive_leader(ACC_EMPTY(acc));
                        }
                        if (live) {
#pragma unroll
                            for (int i = 0; i < 32; ++i) {
                                const int col = 2 * i;
                                const float2 t = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&res[h * 32 + i]));
                                const float f0 = __uint_as_float(v[col]) + (sbias[h * 64 + col] + t.x);
                                const float f1 = __uint_as_float(v[col + 1]) + (sbias[h * 64 + col + 1] + t.y);
                                // the heads see the same bf16-rounded activations as the unfused path
                                const float2 a = __bfloat1622float2(__floats2bfloat162_rn(fmaxf(f0, 0.f), fmaxf(f1, 0.f)));
                                const int cc = h * 64 + col;
                                hd0 = fmaf(a.x, head.w[0][cc], fmaf(a.y, head.w[0][cc + 1], hd0));
                                hd1 = fmaf(a.x, head.w[1][cc], fmaf(a.y, head.w[1][cc + 1], hd1));
                                hd2 = fmaf(a.x, head.w[2][cc], fmaf(a.y, head.w[2][cc + 1], hd2));
                            }
                        }
                        if (h == 0 && C > 64) { hs0 = hd0; hs1 = hd1; hs2 = hd2; hd0 = hd1 = hd2 = 0.f; }
                    }
                    if (C > 64) { hd0 = hs0 + hd0; hd1 = hs1 + hd1; hd2 = hs2 + hd2; }
                    if (live) {  // feat[board][plane*64 + px], plane 0/1 policy, 2 value
                        const int y = row >> 4, b = (row >> 3) & 1, x = row & 7;
                        float* fp = feat + ((size_t)tile * 2 + b) * 192 + y * 8 + x;
                        fp[0] = fmaxf(hd0 + head.b[0], 0.f);
                        fp[64] = fmaxf(hd1 + head.b[1], 0.f);
                        fp[128] = fmaxf(hd2 + head.b[2], 0.f);
                    }
                } else {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_leader(ACC_EMPTY(acc));
                }
            } else {
                // residual slice prefetched into registers BEFORE waiting for the accumulator: its HBM/L2
                // latency hides behind the MMAs of this tile instead of extending the epilogue
                uint32_t res[CH / 2];  // CH bf16 = CH/2 words, as CH/16 256-bit loads
                if (residual && live) {
#pragma unroll
                    for (int i = 0; i < CH / 16; ++i) ldg256(residual + off + hh * CH + i * 16, res + i * 8);
                } else {
#pragma unroll
                    for (int i = 0; i < CH / 2; ++i) res[i] = 0u;
                }
                { PROBE_T0(); mbar_wait_cluster(ACC_FULL(acc), (it >> 1) & 1); PROBE_ADD(p_af); }
                tc_fence_after();
                constexpr int CK = CH < 64 ? CH : 64;  // columns per TMEM staging chunk (v[CK])
#pragma unroll
                for (int h = 0; h < CH / CK; ++h) {
                    uint32_t v[CK];
                    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * C + hh * CH + h * CK);
#pragma unroll
                    for (int j = 0; j < CK / 32; ++j) tc_ld32(taddr + j * 32, v + j * 32);
                    tc_wait_ld();
                    if (h == CH / CK - 1) {  // all of this warp's TMEM reads of the accumulator are done
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_leader(ACC_EMPTY(acc));
                    }
                    if (live) {
#pragma unroll
                        for (int c16 = 0; c16 < CK / 16; ++c16) {  // 16 couts = one 32-byte sector of bf16
                            uint32_t o[8];
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const int col = c16 * 16 + 2 * i;
                                const float2 t = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&res[h * (CK / 2) + c16 * 8 + i]));
                                float f0 = __uint_as_float(v[col]), f1 = __uint_as_float(v[col + 1]);
                                if constexpr (!(C == 128 && CIN == 64)) {  // the first layer has its bias in K and no residual
                                    f0 += sbias[hh * CH + h * CK + col] + t.x;
                                    f1 += sbias[hh * CH + h * CK + col + 1] + t.y;
                                }
                                const __nv_bfloat162 ob = __floats2bfloat162_rn(fmaxf(f0, 0.f), fmaxf(f1, 0.f));
                                o[i] = *reinterpret_cast<const uint32_t*>(&ob);
                            }
                            stg256(out + off + hh * CH + h * CK + c16 * 16, o);
                        }
                    }
                }
            }
        }
#ifdef RVS_CONV_PROBE
        if (!HEAD && CIN == 128 && warp == 2 && lane == 0) {
            g_conv_probe[blockIdx.x * 16 + 5] = p_af;
            g_conv_probe[blockIdx.x * 16 + 6] = clock64() - p_e0;
        }
#endif
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();  // the peer's shared memory / TMEM stay valid until the leader's MMAs are done
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(K::TMEM_COLS));
    }
}


// =============================================================================================
// The whole network up to the head planes in ONE persistent kernel (C = 128): first layer, residual tower, fused heads'
// 1x1 convolutions.
//
// A tile is two whole boards, so a 3x3 convolution never reads across tiles, and every layer maps tile slot `it`
// of pair p to the same CTA: a CTA only ever consumes activations it produced itself.  The network therefore needs NO
// grid-wide synchronisation between layers: each CTA pair walks its own tiles through all layers, and the only
// cross-layer dependencies are inside the CTA --
//   * input tile of layer l  <- the CTA's own epilogue stores of layer l-1 (TMA stores from a staging block): per-
//     epilogue-warp counters in shared memory (`done`), published once the bulk group has completed; the TMA producer
//     thread polls them before it loads the tile;
//   * the resident weights of layer l replace those of layer l-1 in place, one 64-channel half at a time: the half
//     read by the kc = 0 MMAs of the previous layer's last tile is reloaded while its kc = 1 MMAs run (WEMPTY /
//     WFULL barriers per half), so the tensor pipe does not drain at a layer boundary.
// Compared with one launch per layer (conv3x3_tc2_kernel) this removes, per layer: the launch gap, the serial weight
// prologue (the next layer's CTA cannot become resident before this one frees its 224 KB of shared memory: ~1800
// cycles waiting for weights + ~1500 for the first activation box, tools/probe_conv.py), and the tail where early
// CTAs idle until the slowest one finishes; and it lets the tiles go DEPTH FIRST through the layers in small groups,
// so that the activations between layers stay in L2 (see `ngrp` below).  The arithmetic (MMA order, epilogue) is the
// per-layer kernel's, so the results are bit-identical to it (tests/test_gpu_net.py::
// test_tower_kernel_matches_per_layer, tools/stress_tower.py).
// =============================================================================================
struct TowerArgs {
    __nv_bfloat16* buf[3];   // x (block input / residual), t (mid), y (block output); roles rotate per block
    const float* bias;       // [n_layers][C]
    float* feat;             // fused heads' output of the LAST layer (when head != 0)
    const int* n_boards_dev;
    int n_tiles;
    int n_layers;            // 2 x blocks
    int head;
    int conv0;               // 1: the network's first layer (64 -> C, input tiles of 64 channels) runs here too, as layer 0
};

__device__ __forceinline__ int posmod(int x, int n) { const int r = x % n; return r < 0 ? r + n : r; }

// x-slots of the tower's activation boxes: 10 = exactly the slots the nine taps read (x = -1..8).  A (y, board) line
// is then 1280 B, not a whole number of swizzle atoms, which the tensor core does not mind (the swizzle acts on absolute
// address bits, the descriptor's stride between 8-row groups is 1280 B); the 15 KB per stage this saves against 16
// slots are what pays for the epilogue's staging buffers.
#ifndef RVS_TOWER_GROUP
#define RVS_TOWER_GROUP 4
#endif
constexpr int kTowerXS = 10;
constexpr int kTowerABytes = 10 * 2 * kTowerXS * 128;  // 25 KB
constexpr int kTowerStaging = 8 * 2048;                // per epilogue warp: 32 rows x 64 B (32 of the 64 channels it converts)
constexpr int kTowerThreads = 320;                     // warp 0 TMA producer, warp 1 MMA issuer, warps 2-9 epilogue
__device__ __forceinline__ uint64_t make_desc_t(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | (1ull << 16) | ((uint64_t)((kTowerXS * 128) >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// 160 registers x 320 threads leave 14 K registers of the SM free: one CTA of the tree step (72 x 128) or of the heads
// kernel (48 x 256) fits beside a resident CTA of this kernel, which is what lets the two half-batches of a pipelined
// search overlap (rvs_net_search_w1)
template <int C>
__global__ void __cluster_dims__(2, 1, 1) __maxnreg__(160)
conv_tower_kernel(const __grid_constant__ CUtensorMap m0, const __grid_constant__ CUtensorMap m1,
                  const __grid_constant__ CUtensorMap m2, const __grid_constant__ CUtensorMap s0,
                  const __grid_constant__ CUtensorMap s1, const __grid_constant__ CUtensorMap s2,
                  const __grid_constant__ CUtensorMap w_map, const __grid_constant__ CUtensorMap mx0,
                  const __grid_constant__ CUtensorMap w0_map, const __grid_constant__ TowerArgs ta,
                  const __grid_constant__ HeadW head) {
    using K = Cfg2<C, C>;
    static_assert(K::KC == 2, "the weight halves are tied to the two stages of the 128-filter configuration");
    constexpr int kABytesX = kTowerABytes;  // (shadows the per-layer kernels' 16-slot box)
    constexpr int kXSlots = kTowerXS;
    int n_tiles = ta.n_tiles;  // compacted leaf batches: re-read from the device by every role after its griddepcontrol.wait
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* gen = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t w_s = base;
    const uint32_t a_s = base + K::W_TILES * K::W_TILE;
    const uint32_t stg_base = a_s + 2 * kABytesX;  // 1024-aligned: 147456 + 2 x 25600
    unsigned char* stg_gen = gen + K::W_TILES * K::W_TILE + 2 * kABytesX;
    unsigned char* tail = gen + K::W_TILES * K::W_TILE + 2 * kABytesX + kTowerStaging;
    uint64_t* bars = reinterpret_cast<uint64_t*>(tail);
    uint32_t* done = reinterpret_cast<uint32_t*>(tail + 128);   // [8] tiles stored so far, per epilogue warp
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tail + 192);
    float* sbias = reinterpret_cast<float*>(tail + 256);  // [3][C]: three layers' biases in rotation (epilogue warps may be in different layers)
    const uint32_t bar0 = smem_u32(bars);
    auto FULL = [&](int s) { return bar0 + 8u * s; };              // leader
    auto EMPTY = [&](int s) { return bar0 + 8u * (2 + s); };       // both CTAs (multicast commit)
    auto WFULL = [&](int kc) { return bar0 + 8u * (4 + kc); };     // leader: weight half kc of the current layer has landed
    auto WEMPTY = [&](int kc) { return bar0 + 8u * (6 + kc); };    // both CTAs: weight half kc of the finished layer is no longer read
    auto ACC_FULL = [&](int a) { return bar0 + 8u * (8 + a); };    // both CTAs
    auto ACC_EMPTY = [&](int a) { return bar0 + 8u * (10 + a); };  // leader, 16 arrivals (8 epilogue warps x 2 CTAs)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
    // Layer L of this launch: with ta.conv0, L = 0 is the network's FIRST convolution (network.py:97): input tiles of 64
    // channels (3 planes + the two constant-one bias channels), one 64-channel chunk and one K = 16 step per tap, no
    // bias / residual in the epilogue; the residual tower's layer l is L = l + has0.
    const int has0 = ta.conv0 ? 1 : 0;
    const int nL = ta.n_layers + has0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < 2; ++s) { mbar_init(FULL(s), 1); mbar_init(EMPTY(s), 1); mbar_init(WFULL(s), 1); mbar_init(WEMPTY(s), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(ACC_FULL(a), 1); mbar_init(ACC_EMPTY(a), 16); }
        for (int i = 0; i < 8; ++i) done[i] = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(K::TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) pdl_launch_dependents();

    // Layer 0's weights do not depend on the previous kernel: the producer thread requests them, then EVERY thread waits
    // for the previous kernel and reads the batch size -- in uniform control flow.  (Read inside the single-lane role
    // branches, the tile count is a divergent value for the compiler, the MMA issuer's descriptor arithmetic leaves the
    // uniform datapath -- 5 R2UR per tcgen05.mma -- and an MMA takes ~80 instead of ~68 cycles to issue.)
    auto load_weights = [&](int L, int kc) {
        if (rank == 0) mbar_expect_tx(WFULL(kc), 2 * 9 * K::W_TILE);  // both CTAs' halves report to the leader
        for (int tap = 0; tap < 9; ++tap) {
            const uint32_t dst = w_s + (tap * K::KC + kc) * K::W_TILE;
            if (has0 && L == 0) tma2_load_2d(&w0_map, WFULL(kc), dst, 0, tap * C + (int)rank * K::NH);
            else tma2_load_2d(&w_map, WFULL(kc), dst, kc * 64, ((L - has0) * 9 + tap) * C + (int)rank * K::NH);
        }
    };
    if (threadIdx.x == 0) {
        load_weights(0, 0);
        if (!has0) load_weights(0, 1);
    }
    pdl_wait();
    if (ta.n_boards_dev) n_tiles = (__ldcg(ta.n_boards_dev) + 1) >> 1;
    // Tiles of this CTA (same in both CTAs of a pair): pair p owns the tile pairs p, p + n_pairs, ...  Pairs do not pad
    // their share to a common length: a pair that is one tile short just finishes earlier -- no MMAs on zero-filled
    // tiles, and in a pipelined search the other half-batch's CTA takes the SM over at once.
    const int tile_pairs = (n_tiles + 1) >> 1;
    const int n = pair < tile_pairs ? (tile_pairs - pair + n_pairs - 1) / n_pairs : 0;
    // DEPTH FIRST over groups of 4..7 tiles: a group goes through ALL layers before the next group starts.  The live
    // activations of the whole grid are then (group size) x 3 buffers x 32 KB x 148 CTAs ~ 60 MB, which stays in the 126 MB
    // L2 -- layer by layer over all 14 tiles of a CTA, every layer wrote and re-read 67 MB (+ 67 MB of residual) and the
    // 5x128 forward moved ~1.5 GB through HBM.  The chip is power limited under tensor load: with the stores switched off
    // the same cycles ran 17 % faster in time (tools/probe_conv.py), i.e. DRAM traffic is paid for in SM clock.  The price
    // is one weight reload (144 KB per CTA from L2, hidden under the group's last tile) per layer and GROUP.
    const int ngrp = (n + RVS_TOWER_GROUP / 2) / RVS_TOWER_GROUP > 0 ? (n + RVS_TOWER_GROUP / 2) / RVS_TOWER_GROUP : 1;
    auto group_first = [&](int gi) { return n * gi / ngrp; };  // group gi = tile slots [group_first(gi), group_first(gi + 1))
    // Tile order inside a group.  A layer may only load a tile that the previous layer has STORED (and published), so the
    // first tiles a layer visits must be ones the previous layer finished early:
    //   * 3..7 tiles (the normal case): every layer walks the group in the SAME order -- the tile a layer starts with was
    //     stored ng - 1 >= 2 tiles before the previous layer ended (L2 locality is no concern for 3 x 96 KB x ng per CTA);
    //   * >= 8 tiles (RVS_TOWER_GROUP raised): consecutive layers walk in opposite directions (most recently written tiles
    //     are still in L2), rotated by 4, and tiles are published two tiles late (rot_of);
    //   * 1..2 tiles (tiny batches): same order, every tile published at once.
    auto rot_of = [](int ng) { return ng >= 8 ? 4 : 0; };
    // tile slot visited at position i of a layer: (sa * i + sb) mod ng
    auto next_order = [&](int& sa, int& sb, int ng, int rot) { if (rot > 0) { sb = posmod(sa * (ng - 1 - rot) + sb, ng); sa = -sa; } };

    if (warp == 0) {
        if (lane == 0) {  // ===== TMA producer (both CTAs): own tiles, own half of the weight rows =====
            long long p_done = 0, p_empty = 0, p_wempty = 0; (void)p_done; (void)p_empty; (void)p_wempty;
            int stage = 0, sph = 0;      // activation stage ring (two stages)
            int wuse0 = 0, wuse1 = 0;    // layer visits that have used weight half 0 / 1 so far (= loads of that half issued)
            uint32_t tcount = 0, lbase = 0, lbase_prev = 0;  // tiles issued so far; at the start of this / the previous layer visit
            for (int gi = 0; gi < ngrp; ++gi) {
            const int sl0 = group_first(gi), ng = group_first(gi + 1) - sl0, rot = rot_of(ng);
            int sa = rot > 0 ? -1 : 1, sb = rot > 0 ? ng - 1 : 0, pa = 0, pb = 0;
            int xi = 0, ti = 1, yi = 2;
            for (int L = 0; L < nL; ++L) {
                lbase_prev = lbase; lbase = tcount;
                const bool first = has0 && L == 0;
                const int l = L - has0;  // tower layer
                const int in_idx = (l & 1) ? ti : xi;
                const CUtensorMap* amap = first ? &mx0 : (in_idx == 0 ? &m0 : (in_idx == 1 ? &m1 : &m2));
                const int kcs = first ? 1 : 2;
                for (int i = 0; i < ng; ++i, ++tcount) {
                    const int slot = posmod(sa * i + sb, ng);
                    const int tile = ((sl0 + slot) * n_pairs + pair) * 2 + (int)rank;  // may be >= n_tiles: TMA zero-fills
                    if (L > 0) {  // this tile was stored by the CTA's own epilogue warps in the previous layer
                        const uint32_t need = lbase_prev + (uint32_t)posmod(pa * (slot - pb), ng) + 1u;
                        PROBE_T0();
                        for (int w = 0; w < 8; ++w) {
                            uint32_t v;
                            do {
                                asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(done + w)) : "memory");
                            } while (v < need);
                        }
                        PROBE_ADD(p_done);
                    }
#pragma unroll
                    for (int kc = 0; kc < 2; ++kc) {
                        if (kc >= kcs) break;
                        { PROBE_T0(); mbar_wait_cluster(EMPTY(stage), sph ^ 1); PROBE_ADD(p_empty); }
                        if (rank == 0) mbar_expect_tx(FULL(stage), 2 * kABytesX);
                        tma2_load_5d(amap, FULL(stage), a_s + stage * kABytesX, kc * 64, -1, 0, -1, tile);
                        if (++stage == 2) { stage = 0; sph ^= 1; }
                        if (i == 0 && (L > 0 || gi > 0)) {
                            // bring in this layer's half kc once the MMAs of the last layer visit that used it are done
                            const int used = kc == 0 ? wuse0 : wuse1;
                            if (used > 0) { PROBE_T0(); mbar_wait_cluster(WEMPTY(kc), (used - 1) & 1); PROBE_ADD(p_wempty); }
                            load_weights(L, kc);
                        }
                    }
                }
                wuse0 += 1;
                if (kcs == 2) wuse1 += 1;
                pa = sa; pb = sb;
                next_order(sa, sb, ng, rot);
                if (!first && (l & 1)) { const int tmp = xi; xi = yi; yi = tmp; }
            }
            }
#ifdef RVS_CONV_PROBE
            g_conv_probe[blockIdx.x * 16 + 0] = p_empty;
            g_conv_probe[blockIdx.x * 16 + 5] = p_done;
            g_conv_probe[blockIdx.x * 16 + 6] = p_wempty;
#endif
        }
    } else if (warp == 1) {
        if (lane == 0 && rank == 0) {  // ===== MMA issuer (leader only) =====
            uint32_t g = 0;  // tiles issued so far: the accumulator ping-pong runs across layers
            long long p_w = 0, p_acc = 0, p_full = 0, p_start = 0; (void)p_w; (void)p_acc; (void)p_full; (void)p_start;
#ifdef RVS_CONV_PROBE
            p_start = clock64();
#endif
            int stage = 0, sph = 0;
            int wuse0 = 0, wuse1 = 0;
            if (n == 0) {  // empty batch (every game of a wave finished): the weight loads requested above must still land
                mbar_wait_cluster(WFULL(0), 0);  // before the pair's shared memory is given up
                if (!has0) mbar_wait_cluster(WFULL(1), 0);
            }
            for (int gi = 0; gi < ngrp; ++gi) {
            const int ng = group_first(gi + 1) - group_first(gi);
            for (int L = 0; L < nL; ++L) {
#ifdef RVS_CONV_PROBE
                if (gi == 0 && L < 47) g_conv_layer[blockIdx.x * 48 + L] = clock64() - p_start;
#endif
                const bool first = has0 && L == 0;
                const int kcs = first ? 1 : 2;
                const int ksteps = first ? 1 : 4;  // first layer: only channels 0..15 of its input tiles are non-zero
                for (int i = 0; i < ng; ++i, ++g) {
                    const int acc = g & 1;
                    { PROBE_T0(); mbar_wait_cluster(ACC_EMPTY(acc), ((g >> 1) & 1) ^ 1); PROBE_ADD(p_acc); }
                    tc_fence_after();
                    const uint32_t d = tmem_base + (uint32_t)(acc * C);
                    uint32_t accum = 0;
#pragma unroll
                    for (int kc = 0; kc < 2; ++kc) {
                        if (kc >= kcs) break;
                        if (i == 0) { PROBE_T0(); mbar_wait_cluster(WFULL(kc), (kc == 0 ? wuse0 : wuse1) & 1); PROBE_ADD(p_w); }
                        { PROBE_T0(); mbar_wait_cluster(FULL(stage), sph); PROBE_ADD(p_full); }
                        tc_fence_after();
                        const uint32_t a0 = a_s + stage * kABytesX;
#pragma unroll
                        for (int dy = 0; dy < 3; ++dy) {
#pragma unroll
                            for (int dx = 0; dx < 3; ++dx) {
                                const uint32_t wt = w_s + ((dy * 3 + dx) * K::KC + kc) * K::W_TILE;
                                const uint32_t at = a0 + (uint32_t)(dy * 2 * kXSlots + dx) * 128u;
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    if (k < ksteps) {
                                        tc2_mma(d, make_desc_t(at + k * 32), make_desc(wt + k * 32), K::IDESC, accum);
                                        accum = 1;
                                    }
                                }
                            }
                        }
                        tc2_commit_mc(EMPTY(stage));
                        if (i == ng - 1) tc2_commit_mc(WEMPTY(kc));
                        if (++stage == 2) { stage = 0; sph ^= 1; }
                    }
                    tc2_commit_mc(ACC_FULL(acc));
                }
                wuse0 += 1;
                if (kcs == 2) wuse1 += 1;
            }
            }
#ifdef RVS_CONV_PROBE
            g_conv_probe[blockIdx.x * 16 + 1] = p_w;
            g_conv_probe[blockIdx.x * 16 + 2] = p_acc;
            g_conv_probe[blockIdx.x * 16 + 3] = p_full;
            g_conv_probe[blockIdx.x * 16 + 4] = clock64() - p_start;
            g_conv_probe[blockIdx.x * 16 + 7] = n * nL;
            if (nL < 48) g_conv_layer[blockIdx.x * 48 + nL] = (clock64() - p_start) / ngrp;  // (per group, roughly)
#endif
        }
    } else {  // ===== epilogue (both CTAs, own tiles) =====
        // TMEM -> registers -> (+bias, +residual, ReLU, bf16) -> swizzled staging block in shared memory -> TMA store.
        // EIGHT warps: warp (q, hh) converts TMEM lane quarter q (its 32 pixel rows) x channel half hh.  One epilogue
        // warp per scheduler ran the ~900 instructions of a tile's conversion at 0.25 IPC (3600 cycles per tile, more
        // than the tile's MMAs need once the layer boundaries are gone: tools/probe_conv.py); two per scheduler
        // interleave.  The rows leave through a staging block and the TMA unit rather than 32-byte global stores at a
        // 256-byte stride (32 L1 wavefronts per warp instruction).
        const int q = warp & 3;             // TMEM lane quarter this warp may read (hardware rule: warp id % 4)
        const int hh = (warp - 2) >> 2;     // channel half: couts [64 hh, 64 hh + 64)
        const int ew = (warp - 2);          // epilogue warp index 0..7
        const int row = q * 32 + lane;
        const uint32_t stg_blk = stg_base + (uint32_t)ew * 2048u;             // 32 rows x 64 B, SWIZZLE_64B
        const uint32_t stg = stg_blk + (uint32_t)lane * 64u;                  // this thread's staging row
        const uint32_t sw = (uint32_t)((lane >> 1) & 3);                      // 16-byte chunk c of row r sits at chunk c ^ ((r >> 1) & 3)
        uint32_t g = 0;
        int lv = 0;  // layer visits so far (bias row rotation)
        bool pending = false;  // TMA stores of the previous tile not yet known complete (and not yet published)
        long long p_af = 0, p_e0 = 0, p_ld = 0, p_ms = 0, p_pub = 0, p_res = 0; (void)p_af; (void)p_e0; (void)p_ld; (void)p_ms; (void)p_pub; (void)p_res;
#ifdef RVS_CONV_PROBE
        p_e0 = clock64();
#endif
        // Tells the producer thread which of this warp's tiles are in memory.  Every tile commits exactly two bulk groups
        // (empty ones when it stores nothing), so "at most 2 * lag groups pending" means "all tiles but the last `lag`
        // have landed".  With enough tiles per CTA the tile order gives the next layer a head start of rot >= 4 tiles,
        // and publishing runs two tiles late: a TMA store then has two tile times to complete and is never waited for
        // (waiting for the previous tile's store at every tile cost the short first-layer tiles ~6000 cycles each).
        bool lag2 = false;
        auto publish = [&](uint32_t tiles_issued, bool flush) {
            if (lane == 0) {
                uint32_t upto = tiles_issued;
                if (lag2 && !flush) {
                    asm volatile("cp.async.bulk.wait_group 4;" ::: "memory");
                    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the staging block itself is free again
                    upto = tiles_issued >= 2 ? tiles_issued - 2 : 0;
                } else {
                    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
                }
                asm volatile("fence.proxy.async.global;" ::: "memory");  // async-proxy writes -> the warp's later generic residual loads
                asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(smem_u32(done + ew)), "r"(upto) : "memory");
            }
            __syncwarp();
        };
        for (int gi = 0; gi < ngrp; ++gi) {
        const int sl0 = group_first(gi), ng = group_first(gi + 1) - sl0, rot = rot_of(ng);
        if (pending && lag2 != (rot >= 4)) { publish(g, true); pending = false; }  // the lag changes with the group size
        lag2 = rot >= 4;
        int sa = rot > 0 ? -1 : 1, sb = rot > 0 ? ng - 1 : 0;
        const bool defer = ng >= 3;  // publish a tile at the top of a later tile (its stores have landed by then)
        int xi = 0, ti = 1, yi = 2;
        for (int L = 0; L < nL; ++L, ++lv) {
            const bool first = has0 && L == 0;
            const int l = L - has0;  // tower layer (-1: the first convolution, whose folded bias rides in its K dimension)
            const bool is_head = ta.head && L == nL - 1;
            const __nv_bfloat16* residual = (!first && (l & 1)) ? ta.buf[xi] : nullptr;
            const int oi = first ? xi : ((l & 1) ? yi : ti);
            const CUtensorMap* omap = oi == 0 ? &s0 : (oi == 1 ? &s1 : &s2);
            float* sbl = sbias + (lv % 3) * C;
            // every epilogue warp writes the whole (identical) bias row: no barrier between the warps is needed, and
            // three rotating rows keep a warp that is a layer ahead off the row a slower warp still reads
            for (int j = lane; j < C; j += 32) sbl[j] = first ? 0.f : ta.bias[l * C + j];
            __syncwarp();
            for (int i = 0; i < ng; ++i, ++g) {
                const int acc = g & 1;
                const int slot = posmod(sa * i + sb, ng);
                const int tile = ((sl0 + slot) * n_pairs + pair) * 2 + (int)rank;
                const size_t off = ((size_t)tile * kTileRows + row) * C;
                const bool live = tile < n_tiles;
#ifdef RVS_CONV_PROBE
                long long _t1 = clock64();
#endif
                if (pending) {  // earlier tiles' stores have had time to land (this also frees the staging block)
                    PROBE_T0();
                    publish(g, false);
                    pending = false;
                    PROBE_ADD(p_pub);
                }
                if (is_head) {
                    // The three head planes (policy x2, value) are dot products over the 128 channels of a row: each warp sums
                    // its 64 channels in ascending order, the hh = 1 warp hands its partial sums to its hh = 0 partner through
                    // its own (idle) staging block, and plane = sum(channels 0..63) + sum(channels 64..127) -- the order the
                    // per-layer kernel uses too, so both paths give identical bits.
                    uint32_t res[32];
                    if (residual && live) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) ldg256_cg(residual + off + hh * 64 + j * 16, res + j * 8);
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) res[j] = 0u;
                    }
                    { PROBE_T0(); mbar_wait_cluster(ACC_FULL(acc), (g >> 1) & 1); PROBE_ADD(p_af); }
                    tc_fence_after();
                    uint32_t v[64];
                    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * C + hh * 64);
                    tc_ld32(taddr, v);
                    tc_ld32(taddr + 32, v + 32);
                    tc_wait_ld();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_leader(ACC_EMPTY(acc));
                    float hd0 = 0.f, hd1 = 0.f, hd2 = 0.f;
                    if (live) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const int col = 2 * j;
                            const float2 t = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&res[j]));
                            const float f0 = __uint_as_float(v[col]) + (sbl[hh * 64 + col] + t.x);
                            const float f1 = __uint_as_float(v[col + 1]) + (sbl[hh * 64 + col + 1] + t.y);
                            // the heads see the same bf16-rounded activations as the unfused path
                            const float2 a = __bfloat1622float2(__floats2bfloat162_rn(fmaxf(f0, 0.f), fmaxf(f1, 0.f)));
                            const int cc = hh * 64 + col;
                            hd0 = fmaf(a.x, head.w[0][cc], fmaf(a.y, head.w[0][cc + 1], hd0));
                            hd1 = fmaf(a.x, head.w[1][cc], fmaf(a.y, head.w[1][cc + 1], hd1));
                            hd2 = fmaf(a.x, head.w[2][cc], fmaf(a.y, head.w[2][cc + 1], hd2));
                        }
                    }
                    // exchange slot: in the hh = 1 warp's own staging block (free: its stores were waited for above), two
                    // slots alternate so that the partner may still read tile g while tile g + 1 is being summed
                    float* xch = reinterpret_cast<float*>(stg_gen + (size_t)(4 + (ew & 3)) * 2048 + (size_t)(g & 1) * 512);
                    if (hh == 1) { xch[lane] = hd0; xch[32 + lane] = hd1; xch[64 + lane] = hd2; }
                    // the two warps of quarter q; constant ids, so that the kernel reserves 5 of the SM's 16 named barriers and
                    // not all of them (a register id did: then no CTA that uses __syncthreads could be resident beside this one)
                    if (q == 0) asm volatile("bar.sync 1, 64;" ::: "memory");
                    else if (q == 1) asm volatile("bar.sync 2, 64;" ::: "memory");
                    else if (q == 2) asm volatile("bar.sync 3, 64;" ::: "memory");
                    else asm volatile("bar.sync 4, 64;" ::: "memory");
                    if (hh == 0 && live) {  // feat[board][plane*64 + px], plane 0/1 policy, 2 value
                        hd0 += xch[lane]; hd1 += xch[32 + lane]; hd2 += xch[64 + lane];
                        const int y = row >> 4, b = (row >> 3) & 1, x = row & 7;
                        float* fp = ta.feat + ((size_t)tile * 2 + b) * 192 + y * 8 + x;
                        fp[0] = fmaxf(hd0 + head.b[0], 0.f);
                        fp[64] = fmaxf(hd1 + head.b[1], 0.f);
                        fp[128] = fmaxf(hd2 + head.b[2], 0.f);
                    }
                    if (lane == 0) { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
                    if (defer) pending = true;
                    continue;
                }
                uint32_t res[32];  // this warp's 64 channels of the residual row
                if (residual && live) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) ldg256_cg(residual + off + hh * 64 + j * 16, res + j * 8);
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) res[j] = 0u;
                }
#ifdef RVS_CONV_PROBE
                p_res += clock64() - _t1;
#endif
                { PROBE_T0(); mbar_wait_cluster(ACC_FULL(acc), (g >> 1) & 1); PROBE_ADD(p_af); }
                tc_fence_after();
                uint32_t v[64];
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * C + hh * 64);
#ifdef RVS_CONV_PROBE
                long long _t2 = clock64();
#endif
                tc_ld32(taddr, v);
                tc_ld32(taddr + 32, v + 32);
                tc_wait_ld();
                tc_fence_before();  // all of this warp's TMEM reads of the accumulator are done
                __syncwarp();
                if (lane == 0) mbar_arrive_leader(ACC_EMPTY(acc));
#ifdef RVS_CONV_PROBE
                { const long long _t3 = clock64(); p_ld += _t3 - _t2; p_ms -= _t3; }
#endif
                if (live) {
#pragma unroll
                    for (int r = 0; r < 2; ++r) {  // 32 channels per round through the 2-KB staging block
                        uint32_t o[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const int col = r * 32 + 2 * j;
                            const float2 t = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&res[r * 16 + j]));
                            const float f0 = __uint_as_float(v[col]) + (sbl[hh * 64 + col] + t.x);
                            const float f1 = __uint_as_float(v[col + 1]) + (sbl[hh * 64 + col + 1] + t.y);
                            const __nv_bfloat162 ob = __floats2bfloat162_rn(fmaxf(f0, 0.f), fmaxf(f1, 0.f));
                            o[j] = *reinterpret_cast<const uint32_t*>(&ob);
                        }
                        if (r > 0) {  // the TMA unit must have read the first round out of the staging block
                            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                            __syncwarp();
                        }
#pragma unroll
                        for (int c = 0; c < 4; ++c)
                            sts128(stg + (((uint32_t)c ^ sw) << 4), o[4 * c], o[4 * c + 1], o[4 * c + 2], o[4 * c + 3]);
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy STS -> async-proxy read by the TMA store
                        __syncwarp();
                        if (lane == 0) {
                            tma_store_2d(omap, stg_blk, hh * 64 + r * 32, tile * kTileRows + q * 32);
                            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                        }
                    }
                } else if (lane == 0) {  // keep the group count per tile fixed (publish)
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
#ifdef RVS_CONV_PROBE
                p_ms += clock64();
#endif
                if (defer) pending = true;        // published at the top of a later tile: by then the stores have landed
                else publish(g + 1, true);        // one or two tiles per group: the next layer needs this tile at once
            }
            next_order(sa, sb, ng, rot);
            if (!first && (l & 1)) { const int tmp = xi; xi = yi; yi = tmp; }
        }
        }
        publish(g, true);  // nothing may be in flight when the CTA exits
#ifdef RVS_CONV_PROBE
        if (warp == 2 && lane == 0) {
            g_conv_probe[blockIdx.x * 16 + 8] = 0;
            g_conv_probe[blockIdx.x * 16 + 9] = 0;
            g_conv_probe[blockIdx.x * 16 + 10] = p_af;
            g_conv_probe[blockIdx.x * 16 + 11] = clock64() - p_e0;
            g_conv_probe[blockIdx.x * 16 + 12] = p_ld;
            g_conv_probe[blockIdx.x * 16 + 13] = p_ms;
            g_conv_probe[blockIdx.x * 16 + 14] = p_pub;
            g_conv_probe[blockIdx.x * 16 + 15] = p_res;
        }
#endif
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();  // the peer's shared memory / TMEM stay valid until the leader's MMAs are done
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(K::TMEM_COLS));
    }
}

// =============================================================================================
// 256-filter variant (BASELINE config 4: 20 blocks x 256 filters): the 1.18 MB of one layer's
// weights cannot stay resident, so they STREAM through the same mbarrier ring as the activations.
// A CTA pair computes D[256 px, 256 couts]; a stage is one (dx, 64-channel chunk): the CTA's
// activation box (20 KB) + its half (128 couts) of the three vertical taps' weight tiles
// (3 x 16 KB), consumed by 12 MMAs of M256 x N256 x K16 (128 tensor cycles each: 8 KB of operands
// per CTA per MMA = 64 B/cycle of shared memory, below the 128 B/cycle port).  Two 256-column
// accumulators fill the 512 TMEM columns, so the epilogue of tile i still overlaps the MMAs of
// tile i+1.  L2 -> SM traffic is 68 KB per 1536 tensor cycles per SM (44 B/cycle, at the measured
// ~42 B/cycle/SM L2 ceiling): this layer shape is L2-bandwidth bound, not tensor bound.
// =============================================================================================
struct Cfg2S {
    static constexpr int C = 256;
    static constexpr int KC = C / 64;
    static constexpr int NH = C / 2;                  // weight rows (couts) held by each CTA
    static constexpr int W_TILE = NH * 128;           // 16 KB: one (tap, kc) weight tile per CTA
    static constexpr int STAGE = kABytes + 3 * W_TILE;  // 68 KB
    static constexpr int STAGES = 3;
    static constexpr int TMEM_COLS = 2 * C;           // 512: the whole tensor memory
    static constexpr int SMEM = STAGES * STAGE + 1024 /*align*/ + 1280 /*barriers, bias*/;
    static constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(C >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
conv3x3_tc2s_kernel(const __grid_constant__ CUtensorMap a_map, const __grid_constant__ CUtensorMap w_map,
                    const __nv_bfloat16* __restrict__ residual, __nv_bfloat16* __restrict__ out,
                    const float* __restrict__ bias, int n_tiles_arg, const int* __restrict__ n_boards_dev, int rev) {
    using K = Cfg2S;
    // compacted leaf batches: the number of boards is only known on the device (written at least two kernels upstream:
    // this kernel is never the first after the tree step, so it is visible even when this launch overlaps its predecessor's tail)
    const int n_tiles = n_boards_dev ? (*n_boards_dev + 1) >> 1 : n_tiles_arg;
    constexpr int C = K::C;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* gen = smem_raw + (base - smem_u32(smem_raw));
    unsigned char* tail = gen + K::STAGES * K::STAGE;
    uint64_t* bars = reinterpret_cast<uint64_t*>(tail);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tail + 192);
    float* sbias = reinterpret_cast<float*>(tail + 256);  // [256]
    const uint32_t bar0 = smem_u32(bars);
    auto FULL = [&](int s) { return bar0 + 8u * s; };                              // leader
    auto EMPTY = [&](int s) { return bar0 + 8u * (K::STAGES + s); };               // both CTAs (multicast commit)
    auto ACC_FULL = [&](int a) { return bar0 + 8u * (2 * K::STAGES + a); };        // both CTAs
    auto ACC_EMPTY = [&](int a) { return bar0 + 8u * (2 * K::STAGES + 2 + a); };   // leader, 8 arrivals

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
    const int n_iters = (n_tiles + 2 * n_pairs - 1) / (2 * n_pairs);
    auto wait_for_inputs = [&]() { pdl_wait(); };

    if (threadIdx.x == 0) {
        for (int s = 0; s < K::STAGES; ++s) { mbar_init(FULL(s), 1); mbar_init(EMPTY(s), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(ACC_FULL(a), 1); mbar_init(ACC_EMPTY(a), 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < C; i += kThreads) sbias[i] = bias[i];
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(K::TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) pdl_launch_dependents();

    if (warp == 0) {
        if (lane == 0) {  // ===== TMA producer (both CTAs): own activation tile + own half of the weights =====
            wait_for_inputs();
            int stage = 0, phase = 0;
            for (int it = 0; it < n_iters; ++it) {
                const int tile = ((rev ? n_iters - 1 - it : it) * n_pairs + pair) * 2 + (int)rank;  // may be >= n_tiles: TMA zero-fills
                for (int dx = 0; dx < 3; ++dx)
                    for (int kc = 0; kc < K::KC; ++kc) {
                        mbar_wait_cluster(EMPTY(stage), phase ^ 1);
                        if (rank == 0) mbar_expect_tx(FULL(stage), 2 * K::STAGE);
                        const uint32_t st = base + stage * K::STAGE;
                        tma2_load_5d(&a_map, FULL(stage), st, kc * 64, dx - 1, 0, -1, tile);
#pragma unroll
                        for (int dy = 0; dy < 3; ++dy)
                            tma2_load_2d(&w_map, FULL(stage), st + kABytes + dy * K::W_TILE, kc * 64, (dy * 3 + dx) * C + (int)rank * K::NH);
                        if (++stage == K::STAGES) { stage = 0; phase ^= 1; }
                    }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && rank == 0) {  // ===== MMA issuer (leader only) =====
            wait_for_inputs();
            int stage = 0, phase = 0;
            for (int it = 0; it < n_iters; ++it) {
                const int acc = it & 1;
                mbar_wait_cluster(ACC_EMPTY(acc), ((it >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d = tmem_base + (uint32_t)(acc * C);
                uint32_t accum = 0;
                for (int s12 = 0; s12 < 3 * K::KC; ++s12) {
                    mbar_wait_cluster(FULL(stage), phase);
                    tc_fence_after();
                    const uint32_t a0 = base + stage * K::STAGE;
#pragma unroll
                    for (int dy = 0; dy < 3; ++dy) {
                        const uint32_t wt = a0 + kABytes + dy * K::W_TILE;
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            tc2_mma(d, make_desc(a0 + dy * 2048 + k * 32), make_desc(wt + k * 32), K::IDESC, accum);
                            accum = 1;
                        }
                    }
                    tc2_commit_mc(EMPTY(stage));
                    if (++stage == K::STAGES) { stage = 0; phase ^= 1; }
                }
                tc2_commit_mc(ACC_FULL(acc));
            }
        }
    } else {  // ===== epilogue (both CTAs, own tile): 4 chunks of 64 couts, residual chunk h+1 in flight during chunk h =====
        const int q = warp & 3;
        const int row = q * 32 + lane;
        wait_for_inputs();
        for (int it = 0; it < n_iters; ++it) {
            const int acc = it & 1;
            const int tile = ((rev ? n_iters - 1 - it : it) * n_pairs + pair) * 2 + (int)rank;
            const size_t off = ((size_t)tile * kTileRows + row) * C;
            const bool live = tile < n_tiles;
            const bool has_res = residual != nullptr && live;
            uint32_t res[2][32];
#pragma unroll
            for (int i = 0; i < 32; ++i) res[0][i] = 0u;
            if (has_res) {
#pragma unroll
                for (int i = 0; i < 4; ++i) ldg256(residual + off + i * 16, res[0] + i * 8);
            }
            mbar_wait_cluster(ACC_FULL(acc), (it >> 1) & 1);
            tc_fence_after();
#pragma unroll
            for (int h = 0; h < 4; ++h) {
                if (h < 3) {
                    if (has_res) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) ldg256(residual + off + (h + 1) * 64 + i * 16, res[(h + 1) & 1] + i * 8);
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i) res[(h + 1) & 1][i] = 0u;
                    }
                }
                uint32_t v[64];
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * C + h * 64);
                tc_ld32(taddr, v);
                tc_ld32(taddr + 32, v + 32);
                tc_wait_ld();
                if (h == 3) {  // all TMEM reads of this accumulator are done
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_leader(ACC_EMPTY(acc));
                }
                if (live) {
#pragma unroll
                    for (int c16 = 0; c16 < 4; ++c16) {
                        uint32_t o[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int col = c16 * 16 + 2 * i;
                            const __nv_bfloat162 r2 = *reinterpret_cast<const __nv_bfloat162*>(&res[h & 1][c16 * 8 + i]);
                            const float2 t = __bfloat1622float2(r2);
                            const float f0 = __uint_as_float(v[col]) + sbias[h * 64 + col] + t.x;
                            const float f1 = __uint_as_float(v[col + 1]) + sbias[h * 64 + col + 1] + t.y;
                            const __nv_bfloat162 ob = __floats2bfloat162_rn(fmaxf(f0, 0.f), fmaxf(f1, 0.f));
                            o[i] = *reinterpret_cast<const uint32_t*>(&ob);
                        }
                        stg256(out + off + h * 64 + c16 * 16, o);
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(K::TMEM_COLS));
    }
}

template <typename Kern>
int launch_pdl(Kern kern, int threads, int grid, int smem, cudaStream_t s, const CUtensorMap& a_map, const CUtensorMap& w_map,
               const __nv_bfloat16* residual, __nv_bfloat16* out, const float* bias, int n_tiles, const HeadW& head,
               float* feat, const int* n_dev, int rev) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    RVS_CUDA(cudaLaunchKernelEx(&cfg, kern, a_map, w_map, residual, out, bias, n_tiles, head, feat, n_dev, rev));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return 0;
}

int launch_pdl_s(int grid, cudaStream_t s, const CUtensorMap& a_map, const CUtensorMap& w_map, const __nv_bfloat16* residual,
                 __nv_bfloat16* out, const float* bias, int n_tiles, const int* n_dev, int rev) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = Cfg2S::SMEM;
    cfg.stream = s;
    cudaLaunchAttribute attrs[1];
    attrs[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attrs[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attrs;
    cfg.numAttrs = 1;
    RVS_CUDA(cudaLaunchKernelEx(&cfg, conv3x3_tc2s_kernel, a_map, w_map, residual, out, bias, n_tiles, n_dev, rev));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return 0;
}

int encode_act_map(CUtensorMap* m, const void* ptr, int C /*channels of this buffer*/, int64_t n_tiles, int x_slots) {
    EncodeFn enc = get_encode();
    if (!enc) return fail(-9, "cuTensorMapEncodeTiled entry point not available");
    // tiled activation layout [tile][y][board][x][c]  (dims innermost first)
    const cuuint64_t dims[5] = {(cuuint64_t)C, 8, 2, 8, (cuuint64_t)n_tiles};
    const cuuint64_t strides[4] = {(cuuint64_t)C * 2, (cuuint64_t)C * 2 * 8, (cuuint64_t)C * 2 * 16, (cuuint64_t)C * 2 * 128};
    const cuuint32_t box[5] = {64, (cuuint32_t)x_slots, 2, 10, 1};  // 8: one box per dx (streamed kernel); 16: one box for all taps
    const cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(ptr), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(-9, "cuTensorMapEncodeTiled(activations) failed: %d", (int)r);
    return 0;
}

}  // namespace

int conv_tc_plan(ConvTcPlan& plan, const __nv_bfloat16* w, int C, int64_t max_batch, int cin) {
    plan.valid = false;
    plan.C = C;
    plan.max_batch = max_batch;
    if (cin <= 0) cin = C;
    if (!((C == 64 && cin == 64) || (C == 128 && (cin == 128 || cin == 64)) || (C == 256 && cin == 256)))
        return fail(-8, "tcgen05 convolution: unsupported shape %d -> %d", cin, C);
    EncodeFn enc = get_encode();
    if (!enc) return fail(-9, "cuTensorMapEncodeTiled entry point not available");
    Impl* im = plan.impl ? static_cast<Impl*>(plan.impl) : new Impl();
    plan.impl = im;
    im->n_act = 0;
    im->cin = cin;
    // weights [9*C rows (tap, cout)][cin] bf16, box = 64 cin x C/2 couts (the CTA's half of one tap)
    const cuuint64_t dims[2] = {(cuuint64_t)cin, (cuuint64_t)9 * C};
    const cuuint64_t strides[1] = {(cuuint64_t)cin * 2};
    const cuuint32_t es[2] = {1, 1};
    const cuuint32_t box2[2] = {64, (cuuint32_t)(C / 2)};
    CUresult r = enc(&im->w_map2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(w), dims, strides, box2, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(-9, "cuTensorMapEncodeTiled(weights) failed: %d", (int)r);
    // opt in to > 48 KB of dynamic shared memory on THIS device (function attributes are per device, and a
    // process may hold engines on several GPUs, so this is done per plan rather than once per process)
    if (C == 256) {
        RVS_CUDA(cudaFuncSetAttribute(conv3x3_tc2s_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg2S::SMEM));
    } else if (C == 64) {
        RVS_CUDA(cudaFuncSetAttribute(conv3x3_tc2_kernel<64, 64, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg2<64, 64>::SMEM));
        RVS_CUDA(cudaFuncSetAttribute(conv3x3_tc2_kernel<64, 64, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg2<64, 64>::SMEM));
    } else if (cin == 64) {
        RVS_CUDA(cudaFuncSetAttribute(conv3x3_tc2_kernel<128, 64, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg2<128, 64>::SMEM));
        RVS_CUDA(cudaFuncSetAttribute(conv3x3_tc2_kernel<128, 64, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg2<128, 64>::SMEM));
    } else {
        RVS_CUDA(cudaFuncSetAttribute(conv3x3_tc2_kernel<128, 128, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg2<128, 128>::SMEM));
        RVS_CUDA(cudaFuncSetAttribute(conv3x3_tc2_kernel<128, 128, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg2<128, 128>::SMEM));
    }
    plan.valid = true;
    return 0;
}

int conv_tc_launch(const ConvTcPlan& plan, const __nv_bfloat16* in, const __nv_bfloat16* residual, __nv_bfloat16* out,
                   const float* bias, int64_t B, cudaStream_t s, const ConvHeadW* head, float* feat, const int* n_dev,
                   int max_ctas, int64_t cap_boards, int rev, int count_is_fresh) {
    if (!plan.valid || !plan.impl) return fail(-8, "tcgen05 convolution: no plan");
    Impl* im = static_cast<Impl*>(plan.impl);
    static const ConvHeadW zero_head = {};
    if (!(feat && head)) { feat = nullptr; head = &zero_head; }
    const int64_t cap = cap_boards > 0 ? cap_boards : plan.max_batch;  // boards addressable from `in`
    if (B > cap) return fail(-8, "tcgen05 convolution: batch %lld exceeds the buffer capacity %lld", (long long)B, (long long)cap);
    int slot = -1;
    for (int i = 0; i < im->n_act; ++i)
        if (im->act_ptr[i] == in && im->act_cap[i] == cap) slot = i;
    if (slot < 0) {
        if (im->n_act == Impl::kActSlots) im->n_act = 0;
        slot = im->n_act++;
        int rc = encode_act_map(&im->act_map[slot], in, im->cin, (cap + 1) / 2, plan.C == 256 ? 8 : kXSlots);  // tiles beyond the map are zero-filled by TMA
        if (rc) return rc;
        im->act_ptr[slot] = in;
        im->act_cap[slot] = cap;
    }
    const int n_tiles = (int)((B + 1) / 2);
    const int C = plan.C;
    // CTA pairs: grid = 2 x pairs, at most one CTA per SM (max_ctas < 148 leaves SMs to concurrent tree kernels)
    int pair_cap = (max_ctas > 0 && max_ctas < kNumSMs ? max_ctas : kNumSMs) / 2;
    if (pair_cap < 1) pair_cap = 1;
    int pairs = (n_tiles + 1) / 2;
    if (pairs > pair_cap) pairs = pair_cap;
    const CUtensorMap& am = im->act_map[slot];
    if (C == 256) return launch_pdl_s(2 * pairs, s, am, im->w_map2, residual, out, bias, n_tiles, n_dev, rev);
    if (C == 64) {
        if (feat) return launch_pdl(conv3x3_tc2_kernel<64, 64, true, false>, Cfg2<64, 64>::THREADS, 2 * pairs, Cfg2<64, 64>::SMEM, s, am, im->w_map2, residual, out, bias, n_tiles, *head, feat, n_dev, rev);
        return launch_pdl(conv3x3_tc2_kernel<64, 64, false, false>, Cfg2<64, 64>::THREADS, 2 * pairs, Cfg2<64, 64>::SMEM, s, am, im->w_map2, residual, out, bias, n_tiles, *head, feat, n_dev, rev);
    }
    if (im->cin == 64)  // first layer of a 128-filter tower: 64 (3 used) -> 128
    {
        if (count_is_fresh) return launch_pdl(conv3x3_tc2_kernel<128, 64, false, true>, Cfg2<128, 64>::THREADS, 2 * pairs, Cfg2<128, 64>::SMEM, s, am, im->w_map2, residual, out, bias, n_tiles, *head, nullptr, n_dev, rev);
        return launch_pdl(conv3x3_tc2_kernel<128, 64, false, false>, Cfg2<128, 64>::THREADS, 2 * pairs, Cfg2<128, 64>::SMEM, s, am, im->w_map2, residual, out, bias, n_tiles, *head, nullptr, n_dev, rev);
    }
    if (feat) return launch_pdl(conv3x3_tc2_kernel<128, 128, true, false>, Cfg2<128, 128>::THREADS, 2 * pairs, Cfg2<128, 128>::SMEM, s, am, im->w_map2, residual, out, bias, n_tiles, *head, feat, n_dev, rev);
    return launch_pdl(conv3x3_tc2_kernel<128, 128, false, false>, Cfg2<128, 128>::THREADS, 2 * pairs, Cfg2<128, 128>::SMEM, s, am, im->w_map2, residual, out, bias, n_tiles, *head, feat, n_dev, rev);
}

// ---- persistent tower (conv_tower_kernel) ----------------------------------------------------
namespace {
struct TowerImpl {
    CUtensorMap w_map;      // all layers' folded weights: [n_layers * 9 * C rows][C] bf16, box = 64 cin x C/2 couts
    const void* act_ptr[3] = {};
    int64_t act_cap = -1;
    CUtensorMap act_map[3];   // loads: 5-D boxes (10 y x 2 boards x 10 x-slots x 64 channels)
    CUtensorMap w0_map;       // first layer's folded weights [9 * C rows][64] (optional)
    bool has_w0 = false;
    const void* x0_ptr = nullptr;
    CUtensorMap x0_map;       // first layer's input tiles (64 channels)
    CUtensorMap st_map[3];    // stores: [tile rows][C], box = 32 rows x 32 channels (half of an epilogue warp's block), SWIZZLE_64B
};
constexpr int kTowerSmem = Cfg2<128, 128>::W_TILES * Cfg2<128, 128>::W_TILE + 2 * kTowerABytes + kTowerStaging + 1024 /*align*/ + 256 + 3 * 128 * 4;
static_assert(kTowerSmem <= 232448, "persistent tower: shared memory over the 227 KB per-CTA limit");
}  // namespace

int conv_tower_plan(ConvTowerPlan& plan, const __nv_bfloat16* w_slab, const float* bias_slab, int C, int n_layers, int64_t max_batch,
                    const __nv_bfloat16* w0) {
    plan.valid = false;
    if (C != 128 || n_layers < 2 || (n_layers & 1)) return 0;  // other widths keep the per-layer kernels
    EncodeFn enc = get_encode();
    if (!enc) return fail(-9, "cuTensorMapEncodeTiled entry point not available");
    TowerImpl* im = plan.impl ? static_cast<TowerImpl*>(plan.impl) : new TowerImpl();
    plan.impl = im;
    im->act_cap = -1;
    const cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)n_layers * 9 * C};
    const cuuint64_t strides[1] = {(cuuint64_t)C * 2};
    const cuuint32_t es[2] = {1, 1};
    const cuuint32_t box2[2] = {64, (cuuint32_t)(C / 2)};
    CUresult r = enc(&im->w_map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(w_slab), dims, strides, box2, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(-9, "cuTensorMapEncodeTiled(tower weights) failed: %d", (int)r);
    im->has_w0 = w0 != nullptr;
    im->x0_ptr = nullptr;
    if (w0) {  // first layer: [9 * C rows (tap, cout)][64 cin], box = 64 cin x C/2 couts
        const cuuint64_t d0[2] = {64, (cuuint64_t)9 * C};
        const cuuint64_t s0[1] = {64 * 2};
        r = enc(&im->w0_map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(w0), d0, s0, box2, es,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(-9, "cuTensorMapEncodeTiled(first-layer weights) failed: %d", (int)r);
    } else {
        im->w0_map = im->w_map;  // unused placeholder (kernel parameters must be valid descriptors)
    }
    RVS_CUDA(cudaFuncSetAttribute(conv_tower_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTowerSmem));
    plan.C = C;
    plan.n_layers = n_layers;
    plan.bias = bias_slab;
    plan.max_batch = max_batch;
    plan.valid = true;
    return 0;
}

int conv_tower_launch(const ConvTowerPlan& plan, __nv_bfloat16* x, __nv_bfloat16* t, __nv_bfloat16* y, int64_t B, cudaStream_t s,
                      const ConvHeadW& head, float* feat, const int* n_dev, int max_ctas, int64_t cap_boards, const __nv_bfloat16* x0) {
    if (!plan.valid || !plan.impl) return fail(-8, "persistent tower: no plan");
    TowerImpl* im = static_cast<TowerImpl*>(plan.impl);
    const int64_t cap = cap_boards > 0 ? cap_boards : plan.max_batch;
    if (B > cap) return fail(-8, "persistent tower: batch %lld exceeds the buffer capacity %lld", (long long)B, (long long)cap);
    if (im->act_ptr[0] != x || im->act_ptr[1] != t || im->act_ptr[2] != y || im->act_cap != cap) {
        const void* ptrs[3] = {x, t, y};
        for (int i = 0; i < 3; ++i) {
            int rc = encode_act_map(&im->act_map[i], ptrs[i], plan.C, (cap + 1) / 2, kTowerXS);
            if (rc) return rc;
            const cuuint64_t dims[2] = {(cuuint64_t)plan.C, (cuuint64_t)((cap + 1) / 2) * kTileRows};
            const cuuint64_t strides[1] = {(cuuint64_t)plan.C * 2};
            const cuuint32_t box[2] = {32, 32};
            const cuuint32_t es[2] = {1, 1};
            CUresult r = get_encode()(&im->st_map[i], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptrs[i]), dims, strides, box, es,
                                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) return fail(-9, "cuTensorMapEncodeTiled(tower stores) failed: %d", (int)r);
            im->act_ptr[i] = ptrs[i];
        }
        im->act_cap = cap;
        im->x0_ptr = nullptr;
    }
    if (x0 && !im->has_w0) return fail(-8, "persistent tower: first-layer input given but the plan has no first-layer weights");
    if (x0 && im->x0_ptr != x0) {
        int rc = encode_act_map(&im->x0_map, x0, 64, (cap + 1) / 2, kTowerXS);
        if (rc) return rc;
        im->x0_ptr = x0;
    }
    const int n_tiles = (int)((B + 1) / 2);
    int pair_cap = (max_ctas > 0 && max_ctas < kNumSMs ? max_ctas : kNumSMs) / 2;
    if (pair_cap < 1) pair_cap = 1;
    int pairs = (n_tiles + 1) / 2;
    if (pairs > pair_cap) pairs = pair_cap;
    TowerArgs ta;
    ta.conv0 = x0 != nullptr;
    ta.buf[0] = x; ta.buf[1] = t; ta.buf[2] = y;
    ta.bias = plan.bias;
    ta.feat = feat;
    ta.n_boards_dev = n_dev;
    ta.n_tiles = n_tiles;
    ta.n_layers = plan.n_layers;
    ta.head = feat != nullptr;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * pairs);
    cfg.blockDim = dim3(kTowerThreads);
    cfg.dynamicSmemBytes = kTowerSmem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    RVS_CUDA(cudaLaunchKernelEx(&cfg, conv_tower_kernel<128>, im->act_map[0], im->act_map[1], im->act_map[2], im->st_map[0], im->st_map[1],
                                im->st_map[2], im->w_map, x0 ? im->x0_map : im->act_map[0], im->w0_map, ta, head));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return 0;
}

void conv_tower_destroy(ConvTowerPlan& plan) {
    if (plan.impl) delete static_cast<TowerImpl*>(plan.impl);
    plan.impl = nullptr;
    plan.valid = false;
}

void conv_tc_destroy(ConvTcPlan& plan) {
    if (plan.impl) delete static_cast<Impl*>(plan.impl);
    plan.impl = nullptr;
    plan.valid = false;
}

#ifdef RVS_CONV_PROBE
extern "C" int rvs_debug_conv_layers(long long* out) {
    return (int)cudaMemcpyFromSymbol(out, g_conv_layer, sizeof(long long) * 148 * 48);
}
extern "C" int rvs_debug_conv_probe(long long* out) {
    return (int)cudaMemcpyFromSymbol(out, g_conv_probe, sizeof(long long) * 148 * 16);
}
#endif

bool conv_tc_can_fuse_head(const ConvTcPlan& plan) {
    return plan.valid && plan.impl && plan.C <= 128 && static_cast<Impl*>(plan.impl)->cin == plan.C;
}

}  // namespace rvs
