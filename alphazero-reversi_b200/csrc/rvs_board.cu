// rvs_board.cu -- K1 kernels (move generation, flips, apply, random playouts, perft), the K3
// leaf-plane encoder, and their C-ABI entry points.  sm_100a only.
//
// All of these are HBM-streaming integer kernels: one thread per position, 16-byte/8-byte
// coalesced loads, grid-stride loops over a grid sized as a multiple of the SM count.
#include "rvs_board.cuh"
#include "rvs_common.cuh"

#include <cuda_bf16.h>

namespace rvs {

thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};
std::mutex g_stage_mu;

namespace {
struct StageSlot { void* p = nullptr; size_t cap = 0; int dev = -1; };
StageSlot g_slots[16];
}  // namespace

int stage_get(int slot, size_t bytes, void** out) {
    int dev = 0;
    RVS_CUDA(cudaGetDevice(&dev));
    StageSlot& s = g_slots[slot];
    if (s.cap < bytes || s.dev != dev) {
        if (s.p) cudaFree(s.p);
        s.p = nullptr; s.cap = 0;
        size_t want = bytes < 4096 ? 4096 : bytes;
        RVS_CUDA(cudaMalloc(&s.p, want));
        s.cap = want; s.dev = dev;
    }
    *out = s.p;
    return 0;
}

// ------------------------------------------------------------------------------ kernels
template <int RULES>
__global__ void __launch_bounds__(256) legal_masks_kernel(const uint64_t* __restrict__ black,
                                                           const uint64_t* __restrict__ white,
                                                           const uint8_t* __restrict__ side,
                                                           uint64_t* __restrict__ out, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t b = black[i], w = white[i];
        out[i] = side[i] == 1 ? legal_moves<RULES>(b, w) : legal_moves<RULES>(w, b);
    }
}

template <int RULES>
__global__ void __launch_bounds__(256) flip_masks_kernel(const uint64_t* __restrict__ black,
                                                          const uint64_t* __restrict__ white,
                                                          const uint8_t* __restrict__ side,
                                                          const uint8_t* __restrict__ move,
                                                          uint64_t* __restrict__ out, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t b = black[i], w = white[i];
        const int mv = move[i];
        uint64_t f = 0;
        if (mv < 64) {
            const uint64_t m = 1ULL << mv;
            f = side[i] == 1 ? flip_mask<RULES>(b, w, m) : flip_mask<RULES>(w, b, m);
        }
        out[i] = f;
    }
}

template <int RULES>
__global__ void __launch_bounds__(256) apply_moves_kernel(uint64_t* __restrict__ black,
                                                           uint64_t* __restrict__ white,
                                                           uint8_t* __restrict__ side,
                                                           uint8_t* __restrict__ flags,
                                                           const uint8_t* __restrict__ move,
                                                           uint8_t* __restrict__ ok,
                                                           uint64_t* __restrict__ next_legal, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        Board b{black[i], white[i], side[i], flags[i]};
        uint64_t nl = 0;
        const bool done = try_move<RULES>(b, move[i], nl);
        if (done) {
            black[i] = b.black; white[i] = b.white; side[i] = b.side; flags[i] = b.flags;
        } else if (next_legal) {
            nl = is_over(b) ? 0ULL : board_legal<RULES>(b);
        }
        if (ok) ok[i] = done ? 1 : 0;
        if (next_legal) next_legal[i] = nl;
    }
}

// One thread plays one whole game in registers (BASELINE config 1).  HBM traffic is the
// 18-byte result per game; the kernel is bound by integer issue, not bandwidth.
template <int RULES>
__global__ void __launch_bounds__(128) random_playouts_kernel(int64_t n, uint64_t seed, uint64_t first_game,
                                                               uint64_t* __restrict__ out_black,
                                                               uint64_t* __restrict__ out_white,
                                                               uint8_t* __restrict__ out_winner,
                                                               uint8_t* __restrict__ out_plies,
                                                               unsigned long long* __restrict__ total_plies) {
    unsigned long long local = 0;
    for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < n;
         g += (int64_t)gridDim.x * blockDim.x) {
        Board b = start_board();
        const int plies = random_playout<RULES>(b, stream_seed(seed, first_game + (uint64_t)g, 0));
        if (out_black) out_black[g] = b.black;
        if (out_white) out_white[g] = b.white;
        if (out_winner) out_winner[g] = (uint8_t)winner_of(b);
        if (out_plies) out_plies[g] = (uint8_t)plies;
        local += (unsigned long long)plies;
    }
    // warp then block reduction, one atomic per block
    for (int o = 16; o; o >>= 1) local += __shfl_down_sync(0xffffffffu, local, o);
    __shared__ unsigned long long wsum[4];
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x == 0 && total_plies) atomicAdd(total_plies, wsum[0] + wsum[1] + wsum[2] + wsum[3]);
}

// ---- perft: level-synchronous frontier expansion, then a per-thread DFS tail -----------
template <int RULES>
__device__ uint64_t perft_dfs(Board b, int depth) {
    // leaf = depth 0 or game over (oracle/rvs_oracle.c: orc_perft)
    struct Frame { Board b; uint64_t lm; };
    Frame st[8];
    int sp = 0;
    uint64_t total = 0;
    if (depth == 0 || is_over(b)) return 1;
    st[0].b = b;
    st[0].lm = board_legal<RULES>(b);
    if (st[0].lm == 0) return 1;
    int rem = depth;  // remaining depth at frame sp
    while (sp >= 0) {
        Frame& f = st[sp];
        if (rem == 1) {  // children are all leaves
            total += (uint64_t)popc64(f.lm);
            --sp; ++rem;
            continue;
        }
        if (f.lm == 0) { --sp; ++rem; continue; }
        const int idx = ctz64(f.lm);
        f.lm &= f.lm - 1;
        Board c = f.b;
        uint64_t nl;
        apply_move<RULES>(c, idx, nl);
        if (is_over(c)) { total += 1; continue; }
        ++sp; --rem;
        st[sp].b = c;
        st[sp].lm = nl;
    }
    return total;
}

template <int RULES>
__global__ void __launch_bounds__(128) perft_expand_kernel(const uint64_t* __restrict__ in_black,
                                                            const uint64_t* __restrict__ in_white,
                                                            const uint8_t* __restrict__ in_side, int64_t n_in,
                                                            uint64_t* __restrict__ out_black,
                                                            uint64_t* __restrict__ out_white,
                                                            uint8_t* __restrict__ out_side,
                                                            unsigned long long* __restrict__ n_out,
                                                            unsigned long long* __restrict__ leaves) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_in;
         i += (int64_t)gridDim.x * blockDim.x) {
        const Board b{in_black[i], in_white[i], in_side[i], 0};
        const uint64_t lm = board_legal<RULES>(b);
        if (lm == 0) { atomicAdd(leaves, 1ULL); continue; }
        // pass 1: how many children survive (are not terminal)
        int live = 0, dead = 0;
        for (uint64_t m = lm; m; m &= m - 1) {
            Board c = b; uint64_t nl;
            apply_move<RULES>(c, ctz64(m), nl);
            if (is_over(c)) ++dead; else ++live;
        }
        if (dead) atomicAdd(leaves, (unsigned long long)dead);
        if (!live) continue;
        unsigned long long at = atomicAdd(n_out, (unsigned long long)live);
        for (uint64_t m = lm; m; m &= m - 1) {
            Board c = b; uint64_t nl;
            apply_move<RULES>(c, ctz64(m), nl);
            if (is_over(c)) continue;
            out_black[at] = c.black; out_white[at] = c.white; out_side[at] = c.side;
            ++at;
        }
    }
}

template <int RULES>
__global__ void __launch_bounds__(128) perft_tail_kernel(const uint64_t* __restrict__ in_black,
                                                          const uint64_t* __restrict__ in_white,
                                                          const uint8_t* __restrict__ in_side, int64_t n_in,
                                                          int depth, unsigned long long* __restrict__ leaves) {
    unsigned long long local = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_in;
         i += (int64_t)gridDim.x * blockDim.x) {
        local += perft_dfs<RULES>(Board{in_black[i], in_white[i], in_side[i], 0}, depth);
    }
    for (int o = 16; o; o >>= 1) local += __shfl_down_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(leaves, local);
}

// ---- K3: canonical planes (game.py:131-162) ----------------------------------------------
// f32 NCHW: one thread per (position, plane-row pair): each thread writes 2 x float4 x ... ;
// simple mapping: 48 threads per position (3 planes x 16 float4), coalesced 768 B per position.
template <int RULES>
__global__ void __launch_bounds__(256) encode_f32_kernel(const uint64_t* __restrict__ black,
                                                          const uint64_t* __restrict__ white,
                                                          const uint8_t* __restrict__ side,
                                                          float4* __restrict__ out, int64_t n) {
    const int64_t total = n * 48;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
         t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = t / 48;
        const int r = (int)(t - i * 48);
        const int plane = r >> 4, q = r & 15;  // q-th float4 of the plane = squares 4q..4q+3
        const uint64_t b = black[i], w = white[i];
        const bool blk = side[i] == 1;
        const uint64_t P = blk ? b : w, O = blk ? w : b;
        uint64_t src = plane == 0 ? P : (plane == 1 ? O : legal_moves<RULES>(P, O));
        const uint32_t nib = (uint32_t)(src >> (4 * q)) & 15u;
        out[t] = make_float4((float)(nib & 1), (float)((nib >> 1) & 1), (float)((nib >> 2) & 1),
                             (float)((nib >> 3) & 1));
    }
}

// bf16 NHWC with 16 channels (3 used): one thread per square -> 32 bytes (2 x uint4)
template <int RULES>
__global__ void __launch_bounds__(256) encode_bf16_kernel(const uint64_t* __restrict__ black,
                                                           const uint64_t* __restrict__ white,
                                                           const uint8_t* __restrict__ side,
                                                           uint4* __restrict__ out, int64_t n) {
    const int64_t total = n * 64;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
         t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = t >> 6;
        const int sq = (int)(t & 63);
        const uint64_t b = black[i], w = white[i];
        const bool blk = side[i] == 1;
        const uint64_t P = blk ? b : w, O = blk ? w : b;
        const uint64_t L = legal_moves<RULES>(P, O);
        const uint32_t one = 0x3F80u;  // bf16(1.0)
        const uint32_t c0 = ((P >> sq) & 1) ? one : 0u, c1 = ((O >> sq) & 1) ? one : 0u,
                       c2 = ((L >> sq) & 1) ? one : 0u;
        out[2 * t] = make_uint4(c0 | (c1 << 16), c2, 0u, 0u);
        out[2 * t + 1] = make_uint4(0u, 0u, 0u, 0u);
    }
}

}  // namespace rvs

// ------------------------------------------------------------------------------ C ABI
using namespace rvs;

extern "C" {

const char* rvs_last_error(void) { return g_err; }
int rvs_version(void) { return 100; }
int64_t rvs_launch_count(void) { return g_launches.load(); }

int rvs_legal_masks(const uint64_t* black, const uint64_t* white, const uint8_t* side, uint64_t* out_mask,
                    int64_t n, int rules, int mem, void* stream) {
    RVS_NORMALISE_MEM(mem);
    if (n < 0 || (n > 0 && (!black || !white || !side || !out_mask))) return fail(-1, "rvs_legal_masks: bad arguments");
    if (n == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    std::unique_lock<std::mutex> lk(g_stage_mu, std::defer_lock);
    if (mem == RVS_MEM_HOST) lk.lock();
    Arg ab, aw, as, ao;
    int rc;
    if ((rc = arg_in(ab, black, n * 8, mem, 0, s)) || (rc = arg_in(aw, white, n * 8, mem, 1, s)) ||
        (rc = arg_in(as, side, n, mem, 2, s)) || (rc = arg_out(ao, out_mask, n * 8, mem, 3)))
        return rc;
    const int grid = grid_for(n, 256);
    if (rules == RVS_RULES_STRICT)
        RVS_LAUNCH(legal_masks_kernel<RULES_STRICT>, grid, 256, 0, s, (const uint64_t*)ab.dev, (const uint64_t*)aw.dev,
                   (const uint8_t*)as.dev, (uint64_t*)ao.dev, n);
    else
        RVS_LAUNCH(legal_masks_kernel<RULES_REF>, grid, 256, 0, s, (const uint64_t*)ab.dev, (const uint64_t*)aw.dev,
                   (const uint8_t*)as.dev, (uint64_t*)ao.dev, n);
    if (mem == RVS_MEM_HOST) {
        if ((rc = arg_back(ao, s))) return rc;
        RVS_CUDA(cudaStreamSynchronize(s));
    }
    return 0;
}

int rvs_flip_masks(const uint64_t* black, const uint64_t* white, const uint8_t* side, const uint8_t* move,
                   uint64_t* out_flip, int64_t n, int rules, int mem, void* stream) {
    RVS_NORMALISE_MEM(mem);
    if (n < 0 || (n > 0 && (!black || !white || !side || !move || !out_flip))) return fail(-1, "rvs_flip_masks: bad arguments");
    if (n == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    std::unique_lock<std::mutex> lk(g_stage_mu, std::defer_lock);
    if (mem == RVS_MEM_HOST) lk.lock();
    Arg ab, aw, as, am, ao;
    int rc;
    if ((rc = arg_in(ab, black, n * 8, mem, 0, s)) || (rc = arg_in(aw, white, n * 8, mem, 1, s)) ||
        (rc = arg_in(as, side, n, mem, 2, s)) || (rc = arg_in(am, move, n, mem, 3, s)) ||
        (rc = arg_out(ao, out_flip, n * 8, mem, 4)))
        return rc;
    const int grid = grid_for(n, 256);
    if (rules == RVS_RULES_STRICT)
        RVS_LAUNCH(flip_masks_kernel<RULES_STRICT>, grid, 256, 0, s, (const uint64_t*)ab.dev, (const uint64_t*)aw.dev,
                   (const uint8_t*)as.dev, (const uint8_t*)am.dev, (uint64_t*)ao.dev, n);
    else
        RVS_LAUNCH(flip_masks_kernel<RULES_REF>, grid, 256, 0, s, (const uint64_t*)ab.dev, (const uint64_t*)aw.dev,
                   (const uint8_t*)as.dev, (const uint8_t*)am.dev, (uint64_t*)ao.dev, n);
    if (mem == RVS_MEM_HOST) {
        if ((rc = arg_back(ao, s))) return rc;
        RVS_CUDA(cudaStreamSynchronize(s));
    }
    return 0;
}

int rvs_apply_moves(uint64_t* black, uint64_t* white, uint8_t* side, uint8_t* flags, const uint8_t* move,
                    uint8_t* ok, uint64_t* out_next_legal, int64_t n, int rules, int mem, void* stream) {
    RVS_NORMALISE_MEM(mem);
    if (n < 0 || (n > 0 && (!black || !white || !side || !flags || !move))) return fail(-1, "rvs_apply_moves: bad arguments");
    if (n == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    std::unique_lock<std::mutex> lk(g_stage_mu, std::defer_lock);
    if (mem == RVS_MEM_HOST) lk.lock();
    Arg ab, aw, as, af, am, ao, an;
    int rc;
    if ((rc = arg_in(ab, black, n * 8, mem, 0, s)) || (rc = arg_in(aw, white, n * 8, mem, 1, s)) ||
        (rc = arg_in(as, side, n, mem, 2, s)) || (rc = arg_in(af, flags, n, mem, 3, s)) ||
        (rc = arg_in(am, move, n, mem, 4, s)) || (rc = arg_out(ao, ok, n, mem, 5)) ||
        (rc = arg_out(an, out_next_legal, n * 8, mem, 6)))
        return rc;
    const int grid = grid_for(n, 256);
    if (rules == RVS_RULES_STRICT)
        RVS_LAUNCH(apply_moves_kernel<RULES_STRICT>, grid, 256, 0, s, (uint64_t*)ab.dev, (uint64_t*)aw.dev, (uint8_t*)as.dev,
                   (uint8_t*)af.dev, (const uint8_t*)am.dev, (uint8_t*)ao.dev, (uint64_t*)an.dev, n);
    else
        RVS_LAUNCH(apply_moves_kernel<RULES_REF>, grid, 256, 0, s, (uint64_t*)ab.dev, (uint64_t*)aw.dev, (uint8_t*)as.dev,
                   (uint8_t*)af.dev, (const uint8_t*)am.dev, (uint8_t*)ao.dev, (uint64_t*)an.dev, n);
    if (mem == RVS_MEM_HOST) {
        if ((rc = arg_back(ab, s)) || (rc = arg_back(aw, s)) || (rc = arg_back(as, s)) || (rc = arg_back(af, s)) ||
            (rc = arg_back(ao, s)) || (rc = arg_back(an, s)))
            return rc;
        RVS_CUDA(cudaStreamSynchronize(s));
    }
    return 0;
}

int rvs_random_playouts(int64_t n_games, uint64_t seed, uint64_t first_game, int rules, uint64_t* out_black,
                        uint64_t* out_white, uint8_t* out_winner, uint8_t* out_plies, int64_t* out_total_plies,
                        int mem, void* stream) {
    RVS_NORMALISE_MEM(mem);
    if (n_games < 0) return fail(-1, "rvs_random_playouts: bad arguments");
    if (out_total_plies) *out_total_plies = 0;
    if (n_games == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    std::unique_lock<std::mutex> lk(g_stage_mu);
    Arg ab, aw, ai, ap;
    int rc;
    void* dtotal = nullptr;
    if ((rc = arg_out(ab, out_black, n_games * 8, mem, 0)) || (rc = arg_out(aw, out_white, n_games * 8, mem, 1)) ||
        (rc = arg_out(ai, out_winner, n_games, mem, 2)) || (rc = arg_out(ap, out_plies, n_games, mem, 3)) ||
        (rc = stage_get(7, 8, &dtotal)))
        return rc;
    RVS_CUDA(cudaMemsetAsync(dtotal, 0, 8, s));
    const int grid = grid_for(n_games, 128, 16);
    if (rules == RVS_RULES_STRICT)
        RVS_LAUNCH(random_playouts_kernel<RULES_STRICT>, grid, 128, 0, s, n_games, seed, first_game, (uint64_t*)ab.dev,
                   (uint64_t*)aw.dev, (uint8_t*)ai.dev, (uint8_t*)ap.dev, (unsigned long long*)dtotal);
    else
        RVS_LAUNCH(random_playouts_kernel<RULES_REF>, grid, 128, 0, s, n_games, seed, first_game, (uint64_t*)ab.dev,
                   (uint64_t*)aw.dev, (uint8_t*)ai.dev, (uint8_t*)ap.dev, (unsigned long long*)dtotal);
    if (mem == RVS_MEM_HOST) {
        if ((rc = arg_back(ab, s)) || (rc = arg_back(aw, s)) || (rc = arg_back(ai, s)) || (rc = arg_back(ap, s))) return rc;
    }
    if (out_total_plies) {
        RVS_CUDA(cudaMemcpyAsync(out_total_plies, dtotal, 8, cudaMemcpyDeviceToHost, s));
        RVS_CUDA(cudaStreamSynchronize(s));
    } else if (mem == RVS_MEM_HOST) {
        RVS_CUDA(cudaStreamSynchronize(s));
    }
    return 0;
}

int rvs_perft(uint64_t black, uint64_t white, int side, int depth, int rules, uint64_t* out_count, void* stream) {
    if (!out_count || depth < 0 || depth > 16 || (side != 1 && side != 2) || (black & white))
        return fail(-1, "rvs_perft: bad arguments");
    cudaStream_t s = (cudaStream_t)stream;
    if (depth == 0) { *out_count = 1; return 0; }
    const int tail = depth < 4 ? depth : 4;
    const int levels = depth - tail;
    unsigned long long* dcnt = nullptr;  // [0] leaves, [1] n_out
    RVS_CUDA(cudaMalloc(&dcnt, 16));
    RVS_CUDA(cudaMemsetAsync(dcnt, 0, 16, s));
    uint64_t *ib = nullptr, *iw = nullptr; uint8_t* is = nullptr;
    int64_t n_in = 1;
    RVS_CUDA(cudaMalloc(&ib, 8)); RVS_CUDA(cudaMalloc(&iw, 8)); RVS_CUDA(cudaMalloc(&is, 1));
    const uint8_t side8 = (uint8_t)side;
    RVS_CUDA(cudaMemcpyAsync(ib, &black, 8, cudaMemcpyHostToDevice, s));
    RVS_CUDA(cudaMemcpyAsync(iw, &white, 8, cudaMemcpyHostToDevice, s));
    RVS_CUDA(cudaMemcpyAsync(is, &side8, 1, cudaMemcpyHostToDevice, s));
    int rc = 0;
    for (int l = 0; l < levels && n_in > 0; ++l) {
        const int64_t cap = n_in * 33;
        if (cap * 17 > (int64_t)16 << 30) { rc = fail(-2, "rvs_perft: frontier too large at level %d", l); break; }
        uint64_t *ob = nullptr, *ow = nullptr; uint8_t* os = nullptr;
        if (cudaMalloc(&ob, cap * 8) != cudaSuccess || cudaMalloc(&ow, cap * 8) != cudaSuccess ||
            cudaMalloc(&os, cap) != cudaSuccess) {
            cudaFree(ob); cudaFree(ow); cudaFree(os);
            rc = fail(-3, "rvs_perft: out of memory at level %d", l);
            break;
        }
        cudaMemsetAsync(dcnt + 1, 0, 8, s);
        const int grid = grid_for(n_in, 128, 16);
        if (rules == RVS_RULES_STRICT)
            perft_expand_kernel<RULES_STRICT><<<grid, 128, 0, s>>>(ib, iw, is, n_in, ob, ow, os, dcnt + 1, dcnt);
        else
            perft_expand_kernel<RULES_REF><<<grid, 128, 0, s>>>(ib, iw, is, n_in, ob, ow, os, dcnt + 1, dcnt);
        g_launches.fetch_add(1);
        unsigned long long n_out = 0;
        cudaMemcpyAsync(&n_out, dcnt + 1, 8, cudaMemcpyDeviceToHost, s);
        cudaError_t e = cudaStreamSynchronize(s);
        cudaFree(ib); cudaFree(iw); cudaFree(is);
        ib = ob; iw = ow; is = os;
        n_in = (int64_t)n_out;
        if (e != cudaSuccess) { rc = fail(-100 - (int)e, "rvs_perft: %s", cudaGetErrorString(e)); break; }
    }
    if (rc == 0 && n_in > 0) {
        const int grid = grid_for(n_in, 128, 16);
        if (rules == RVS_RULES_STRICT)
            perft_tail_kernel<RULES_STRICT><<<grid, 128, 0, s>>>(ib, iw, is, n_in, tail, dcnt);
        else
            perft_tail_kernel<RULES_REF><<<grid, 128, 0, s>>>(ib, iw, is, n_in, tail, dcnt);
        g_launches.fetch_add(1);
    }
    unsigned long long total = 0;
    if (rc == 0) {
        cudaMemcpyAsync(&total, dcnt, 8, cudaMemcpyDeviceToHost, s);
        cudaError_t e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) rc = fail(-100 - (int)e, "rvs_perft: %s", cudaGetErrorString(e));
    }
    cudaFree(ib); cudaFree(iw); cudaFree(is); cudaFree(dcnt);
    *out_count = total;
    return rc;
}

int rvs_encode_planes(const uint64_t* black, const uint64_t* white, const uint8_t* side, void* out, int64_t n,
                      int layout, int rules, int mem, void* stream) {
    RVS_NORMALISE_MEM(mem);
    if (n < 0 || (n > 0 && (!black || !white || !side || !out))) return fail(-1, "rvs_encode_planes: bad arguments");
    if (layout != RVS_PLANES_F32_NCHW && layout != RVS_PLANES_BF16_NHWC) return fail(-1, "rvs_encode_planes: bad layout");
    if (n == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    std::unique_lock<std::mutex> lk(g_stage_mu, std::defer_lock);
    if (mem == RVS_MEM_HOST) lk.lock();
    const size_t per = layout == RVS_PLANES_F32_NCHW ? 768 : 64 * 16 * 2;
    Arg ab, aw, as, ao;
    int rc;
    if ((rc = arg_in(ab, black, n * 8, mem, 0, s)) || (rc = arg_in(aw, white, n * 8, mem, 1, s)) ||
        (rc = arg_in(as, side, n, mem, 2, s)) || (rc = arg_out(ao, out, n * per, mem, 3)))
        return rc;
    if (layout == RVS_PLANES_F32_NCHW) {
        const int grid = grid_for(n * 48, 256);
        if (rules == RVS_RULES_STRICT)
            RVS_LAUNCH(encode_f32_kernel<RULES_STRICT>, grid, 256, 0, s, (const uint64_t*)ab.dev, (const uint64_t*)aw.dev,
                       (const uint8_t*)as.dev, (float4*)ao.dev, n);
        else
            RVS_LAUNCH(encode_f32_kernel<RULES_REF>, grid, 256, 0, s, (const uint64_t*)ab.dev, (const uint64_t*)aw.dev,
                       (const uint8_t*)as.dev, (float4*)ao.dev, n);
    } else {
        const int grid = grid_for(n * 64, 256);
        if (rules == RVS_RULES_STRICT)
            RVS_LAUNCH(encode_bf16_kernel<RULES_STRICT>, grid, 256, 0, s, (const uint64_t*)ab.dev, (const uint64_t*)aw.dev,
                       (const uint8_t*)as.dev, (uint4*)ao.dev, n);
        else
            RVS_LAUNCH(encode_bf16_kernel<RULES_REF>, grid, 256, 0, s, (const uint64_t*)ab.dev, (const uint64_t*)aw.dev,
                       (const uint8_t*)as.dev, (uint4*)ao.dev, n);
    }
    if (mem == RVS_MEM_HOST) {
        if ((rc = arg_back(ao, s))) return rc;
        RVS_CUDA(cudaStreamSynchronize(s));
    }
    return 0;
}

}  // extern "C"
