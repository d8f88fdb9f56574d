/*
 * snaq_engine.cu — CUDA kernels (sm_100a) and the C ABI of include/snaq_b200.h.
 *
 * Device-resident state per context (one context <=> one GPU + one stream):
 *   taxa   [nq][4] u16      quartet taxon ids                      (set_data)
 *   obs    [nq][3] f64      observed CFs                           (set_data / set_obscf)
 *   sampled[nq]    u8       q.sampled                              (set_sampled)
 *   off    [nq+1]  u32  \   CSR table of quartet programs, sorted by (class, program / term hash): the flattened
 *   prog   [words] u16  /   replacement of the reference's per-quartet QuartetNetwork objects (set_network)
 *   W      [nq][4] f64      100*obs in program order + sum 100*obs*log(obs);  WA [nA12][2]: the pairs of the short
 *                           additive programs
 *   (SNAQ_OPT_DEDUP) off_u / prog_u / W_u: the same table with one entry per DISTINCT program, weights summed
 *
 * Kernels:
 *   k_build_count / k_build_emit : extractQuartet! for every quartet (thread per quartet), sort keys
 *   k_dedup_* / k_weights_u      : runs of identical programs -> compressed table, compensated weight sums
 *   k_weights                    : obs -> W, WA, class-A constant
 *   k_eval_lanes / k_eval_lanes_p: the batched kernels.  lane = candidate parameter vector, a warp walks a contiguous
 *                                  share of a chunk of quartets; the quartet program is warp-uniform so there is no
 *                                  divergence, x is staged transposed in shared memory so the per-slot gathers are
 *                                  conflict-free, t1/exp/log and hybridisation terms are reused while the (sorted)
 *                                  programs repeat, sums stay thread-private until a fixed-order block reduction.
 *                                  _p: persistent blocks, xf tile resident, chunks through two TMA stages (big tiles)
 *   k_eval_direct                : thread per quartet for small batches (P < 16; HBM-bound at P=1): warp-private TMA
 *                                  rings; also the per-quartet read-back (expCF, logPseudoLik, deltaCF)
 *   k_eval_grad / _small         : objective + analytic gradient (+ curvature diagonal) in one pass, fixed-point bins.
 *                                  Large tables: class A through a TMA ring in runs of consecutive tiles, per-program
 *                                  memo, grid-wide 64-bit integer bins; small tables: plain loads, per-block rows
 *   k_chunk_sums / k_sample      : deltaCF-weighted quartet draw of the move proposals (inverse-CDF walk)
 *   k_reduce_partials*           : second, fixed-order pass over per-block partial sums
 *   k_update_bl, k_describe, k_hasedge : updateBL! table, parity descriptors
 *
 * Tensor cores are not used: the path is a few exp/log per (quartet, candidate), not a contraction.
 */
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nvtx3/nvToolsExt.h>   // header-only NVTX v3: ranges around build / expand / eval / reduce (no-ops without a profiler)
#include <nccl.h>            // types and prototypes only: the library is resolved at run time (snaq_comm_init), never linked
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/snaq_b200.h"
#include "snaq_core.h"
#include "snaq_hasedge.h"
#include "snaq_tables.h"

#define SNAQ_S_RANGE 8u   /* a parameter outside its bounds reached the device */
#define SNAQ_LCA_MAX_TAXA 4096          // pair table of LCAs: 2*ntaxa^2 bytes (32 MB at the cap), beyond it the builder walks
#define SNAQ_STAGE_WORDS 12   /* program words per quartet kept from the classification pass (3 chunks: all but the longest programs) */
// Small batches from the host travel as a kernel ARGUMENT (no host->device copy call, no staging buffer), and the
// results are written by the last block straight into page-locked host memory (no device->host copy call).
#define XARG_MAX 448      /* doubles (3.5 KB of kernel parameter space) */
struct XArg { double v[XARG_MAX]; };
#define SNAQ_S_DONE 0x80000000u   /* status word in page-locked memory: the kernel has written its results */
#define LANES_MIN_P 16    /* batches of at least this many candidates use the batched kernels */
#define SNAQ_S_COMM 16u   /* quartet-sharded mode: a peer's partial sums did not arrive (fused epilogue of k_eval_direct) */
// Quartet-sharded mode, small batches: the all-reduce of the per-candidate partial sums is FUSED into the epilogue of
// k_eval_direct.  Every rank owns a mailbox array in its own HBM, mapped into every peer through CUDA IPC (NVLink /
// NVSwitch peer memory): box[slot][sender] = {v[4], seq}.  The last block of rank r stores its partial sums and then the
// launch's sequence number into box[seq & 1][r] of EVERY rank (plain stores over NVLink, fence.sys in between), spins
// until the sequence number has arrived in all `world` entries of its own box[seq & 1], and adds the entries in rank
// order: every rank gets the same bits, with one launch and no second kernel.  Two slots, because a fast rank may
// already deliver launch k+1 while a slow one still reads launch k (it cannot deliver k+2 before the slow one has
// delivered k+1, i.e. finished k).
struct PeerMail { double v[4]; unsigned long long seq; unsigned long long pad[3]; };   // 64 bytes
#define SNAQ_MAX_WORLD 16
struct PeerBox {
    PeerMail *peer[SNAQ_MAX_WORLD];   // peer[r]: rank r's box array [2][world] as mapped into this process (peer[rank] = own)
    int rank, world;
    unsigned long long seq;            // 0: not sharded / no fused epilogue for this launch
};
#define RN_TQ 64          /* k_eval_runs: quartets per tree-like tile (64 x 16 B weight pairs) */
#define RN_TQC 32         /*              quartets per 4-cycle tile (32 x 32 B weights)        */

// --------------------------------------------------------------------------------------------
// math tables (copied to shared memory by every evaluation kernel)
// --------------------------------------------------------------------------------------------
__device__ const double g_math_tab[SNAQ_MATH_DOUBLES] = SNAQ_EXP_TAB_INIT_LOGS;   // 2^(j/16) | 1/c_j | -log(1/c_j)

// dst must be 128-byte aligned: each 16-entry table is then one shared-memory row, so lookups with
// arbitrary per-lane indices are bank-conflict free
__device__ __forceinline__ SnaqMath load_math_tables(double *dst, int tid) {
    if (tid < SNAQ_MATH_DOUBLES) dst[tid] = g_math_tab[tid];
    return SnaqMath{dst, dst + 16, dst + 32};
}

// ---- TMA 1-D bulk copies (cp.async.bulk) completing on an mbarrier -------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, unsigned bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// Programmatic dependent launch: a kernel launched with launch_pdl may start while its predecessor in the stream is still
// running; it calls grid_dependency_wait() before the first access to anything the predecessor writes (a no-op when the
// kernel was launched the ordinary way).  Used for the expand -> evaluate -> reduce chain of one batched step, whose two
// launch gaps are 7-14 % of the step.
__device__ __forceinline__ void grid_dependency_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// the dependent kernel may be scheduled once every block of this grid has got here (it still waits for this grid's memory)
__device__ __forceinline__ void grid_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
template <class... KArgs, class... Args>
static cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// --------------------------------------------------------------------------------------------
// descriptor build
// --------------------------------------------------------------------------------------------
__device__ __forceinline__ void load_taxa(const uint16_t *__restrict__ taxa, int64_t q, uint16_t tx[4]) {
    const ushort4 t4 = reinterpret_cast<const ushort4 *>(taxa)[q];
    tx[0] = t4.x; tx[1] = t4.y; tx[2] = t4.z; tx[3] = t4.w;
}

// pass 1 (original quartet order): program length in 4-word chunks and class of every quartet
// SORT KEY: (sort class << 28) | 28-bit hash: of the whole program for class A, of the hybridisation term for classes B
// and C.  Quartets with the same additive program / the same term (same cycle, same hit positions) become neighbours, so
// the batched kernels can reuse what they have just computed (see the m*_ variables of k_eval_lanes).
__device__ __forceinline__ uint32_t term_hash(unsigned a, unsigned b, unsigned c, unsigned d) {
    uint32_t h = a * 0x9E3779B1u;
    h = (h ^ b) * 0x85EBCA77u;
    h = (h ^ c) * 0xC2B2AE3Du;
    h = (h ^ d) * 0x27D4EB2Fu;
    return (h ^ (h >> 15)) & 0x0FFFFFFFu;
}
// blob-tree LCA of every pair of taxa (SnaqTables::leaf_lca): one thread per ordered pair, once per topology
__global__ void __launch_bounds__(256) k_leaf_lca(SnaqTables T, uint16_t *__restrict__ out) {
    const int n = T.ntaxa;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * n) return;
    const int i = idx / n, j = idx - i * n;
    const int u = T.leaf_tnode[i], v = T.leaf_tnode[j];
    out[idx] = (u == 0xFFFF || v == 0xFFFF) ? (uint16_t)0xFFFF : (uint16_t)snaq_lca(T, u, v);
}
// snaq_patch_network: the quartets that hold a flagged taxon, as a dense list (their order does not matter: every result is
// stored by quartet index); the others keep their length, sort key, staged program and header and only enter the
// statistics (stats as in k_build_count)
__global__ void __launch_bounds__(256) k_dirty_list(const uint16_t *__restrict__ taxa, int64_t nq, const uint8_t *__restrict__ dirty_taxon,
                                                    const uint32_t *__restrict__ len4, const uint32_t *__restrict__ cls,
                                                    uint32_t *__restrict__ list, uint32_t *__restrict__ nlist,
                                                    unsigned long long *__restrict__ stats) {
    // a block takes 1024 consecutive quartets (four per thread) and one contiguous piece of the list: one atomic per block
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __shared__ unsigned long long ssum[8];
    __shared__ unsigned smax[8], scnt[5], wcnt[32], wbase[32];
    if (threadIdx.x < 5) scnt[threadIdx.x] = 0;
    __syncthreads();
    unsigned long long s = 0;
    unsigned mx = 0, mk[4];
    bool dirty[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int64_t q = (int64_t)blockIdx.x * 1024 + k * 256 + threadIdx.x;
        dirty[k] = false;
        unsigned n = 0, c = 5;
        if (q < nq) {
            uint16_t tx[4];
            load_taxa(taxa, q, tx);
            dirty[k] = (dirty_taxon[tx[0]] | dirty_taxon[tx[1]] | dirty_taxon[tx[2]] | dirty_taxon[tx[3]]) != 0;
            if (!dirty[k]) { n = len4[q]; c = cls[q] >> 28; }
        }
        mk[k] = __ballot_sync(0xffffffffu, dirty[k]);
        if (lane == 0) wcnt[k * 8 + warp] = (unsigned)__popc(mk[k]);
        if (c < 5) atomicAdd(&scnt[c], 1u);
        s += n;
        mx = max(mx, n);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned t = 0;
        for (int k = 0; k < 32; k++) { wbase[k] = t; t += wcnt[k]; }
        const unsigned b0 = t ? atomicAdd(nlist, t) : 0u;
        for (int k = 0; k < 32; k++) wbase[k] += b0;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; k++)
        if (dirty[k]) list[wbase[k * 8 + warp] + __popc(mk[k] & ((1u << lane) - 1u))] = (uint32_t)((int64_t)blockIdx.x * 1024 + k * 256 + threadIdx.x);
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_down_sync(0xffffffffu, s, o);
        mx = max(mx, __shfl_down_sync(0xffffffffu, mx, o));
    }
    if (lane == 0) { ssum[warp] = s; smax[warp] = mx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        unsigned tm = 0;
        for (int k = 0; k < 8; k++) { t += ssum[k]; tm = max(tm, smax[k]); }
        if (t) atomicAdd(&stats[0], t);
        if (tm) atomicMax(&stats[1], (unsigned long long)tm);
        for (int k = 0; k < 5; k++) if (scnt[k]) atomicAdd(&stats[2 + k], (unsigned long long)scnt[k]);
    }
}

__global__ void __launch_bounds__(128) k_build_count(SnaqTables T, const uint16_t *__restrict__ taxa, int64_t nq, const int dedup,
                                                     uint32_t *__restrict__ len4, uint32_t *__restrict__ cls,
                                                     uint32_t *__restrict__ iota, unsigned long long *__restrict__ stats,
                                                     uint32_t *__restrict__ status, const uint32_t *__restrict__ list,
                                                     const uint32_t *__restrict__ nlist,
                                                     uint16_t *__restrict__ stage, uint8_t *__restrict__ hdr0) {
    // stats: [0] total chunks, [1] max chunks, [2..6] quartets per sort class
    // sort classes: 0 = class A with a one-chunk program, 1 = class A with two chunks, 2 = longer class A,
    //               3 = class B, 4 = class C
    // stage / hdr0 (original quartet order): the first SNAQ_STAGE_WORDS words of the program, NULL padded, and the header
    // byte: k_build_emit copies short programs from there instead of deriving them a second time.
    // list / nlist (snaq_patch_network, k_dirty_list): only the listed quartets are re-described; the others keep the length,
    // sort key, staged program and header of the previous topology (their root paths and blocks have not changed).
    int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (list) {                                                // k_dirty_list's quartets only (the grid covers the whole table)
        const uint32_t nl = *nlist;
        if ((int64_t)blockIdx.x * blockDim.x >= (int64_t)nl) return;
        q = q < (int64_t)nl ? (int64_t)list[q] : nq;
    }
    unsigned n = 0, c = 5;
    if (q < nq) {
        uint16_t tx[4];
        load_taxa(taxa, q, tx);
        {
        uint16_t head[24];                                     // the first words of the program: enough for the term key
        SnaqWriter w{head, 0, 24};
        uint8_t hdr = 0;
        int e = snaq_build_quartet(T, tx, w, &hdr, nullptr);
        if (e) { atomicMax(status, (uint32_t)e); w.n = 1; hdr = 0; }
        n = (unsigned)((w.n + 3) >> 2);
        if (n == 0) n = 1;
        c = SNAQ_HDR_CLASS(hdr);
        c = (c == 0) ? (n == 1 ? 0u : (n == 2 ? 1u : 2u)) : c + 2u;
        uint32_t key = 0;
        if (c == 4u && w.n >= 4) key = term_hash(head[0], head[1], head[2], head[3]);
        else if (c == 3u) {
            const int pos = 1 + 2 * (int)(head[0] & 255u);     // first term header, behind the count word and the plain pairs
            if (pos + 3 < 24 && pos + 3 < w.n) key = (head[pos] & 3u) == SNAQ_T2 ? term_hash(head[pos], head[pos + 1], head[pos + 2], 0u)
                                                   : term_hash(head[pos] & ~4u, head[pos + 1], head[pos + 2], head[pos + 3]);
        }
        const uint32_t term_key = key;
        if (dedup || c <= 3u) {
            // class A (always) and program deduplication (every class): the key is a hash of the WHOLE program, so that
            // identical programs become neighbours (programs longer than the captured head hash only its first 24 words:
            // they may be split into several runs)
            uint32_t h = (uint32_t)w.n * 0x9E3779B1u;
            const int m = w.n < 24 ? w.n : 24;
            for (int i = 0; i < m; i++) h = (h ^ head[i]) * 0x85EBCA77u;
            key = (h ^ (h >> 15)) & 0x0FFFFFFFu;
            // class B: by term, then by program -- also on the run table: programs that share a term stay neighbours and
            // the batched kernels reuse the term's value (k_eval_lanes, m2_ / m4_)
            // (10 bits of term, 18 of program: two programs of one term bucket that share the 18 bits would interleave and
            // split each other's runs -- 30 % more runs at 200 taxa with 14 + 14 bits, < 1 % with 10 + 18)
            if (c == 3u) key = ((term_key & 0x3FFu) << 18) | (key & 0x3FFFFu);
        }
        len4[q] = n;
        cls[q] = (c << 28) | key;
        iota[q] = (uint32_t)q;
        hdr0[q] = e ? 0 : hdr;
        {
            const unsigned nul = (unsigned)T.n_xe;
            uint2 *st = reinterpret_cast<uint2 *>(stage + (size_t)q * SNAQ_STAGE_WORDS);
#pragma unroll
            for (int k = 0; k < SNAQ_STAGE_WORDS / 4; k++) {
                unsigned v[4];
#pragma unroll
                for (int j = 0; j < 4; j++) v[j] = (4 * k + j < w.n && !e) ? (unsigned)head[4 * k + j] : nul;
                st[k] = make_uint2(v[0] | (v[1] << 16), v[2] | (v[3] << 16));
            }
        }
        }
    }
    __shared__ unsigned long long ssum[4];
    __shared__ unsigned smax[4], scnt[5];
    if (threadIdx.x < 5) scnt[threadIdx.x] = 0;
    __syncthreads();
    if (c < 5) atomicAdd(&scnt[c], 1u);
    unsigned long long s = n;
    unsigned m = n;
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_down_sync(0xffffffffu, s, o);
        m = max(m, __shfl_down_sync(0xffffffffu, m, o));
    }
    if ((threadIdx.x & 31) == 0) { ssum[threadIdx.x >> 5] = s; smax[threadIdx.x >> 5] = m; }
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned long long t = ssum[0] + ssum[1] + ssum[2] + ssum[3];
        if (t) atomicAdd(&stats[0], t);
        if (t) atomicMax(&stats[1], (unsigned long long)max(max(smax[0], smax[1]), max(smax[2], smax[3])));
        for (int k = 0; k < 5; k++) if (scnt[k]) atomicAdd(&stats[2 + k], (unsigned long long)scnt[k]);
    }
}

// lengths in class-sorted order (then scanned in place into offsets)
// pad_at: sorted position whose program gets one extra (all-NULL) chunk, or -1: keeps the two-chunk region
// 16-byte aligned for the bulk copies of k_eval_direct
__global__ void __launch_bounds__(256) k_gather_len(const uint32_t *__restrict__ perm, const uint32_t *__restrict__ len4,
                                                    int64_t nq, int64_t pad_at, uint32_t *__restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nq) out[i] = len4[perm[i]] + (i == pad_at ? 1u : 0u);
}

// pass 2 (sorted order): emit the program of quartet perm[i] at chunk offset off[i], NULL-padded
__global__ void __launch_bounds__(128) k_build_emit(SnaqTables T, const uint16_t *__restrict__ taxa, int64_t nq,
                                                    const uint32_t *__restrict__ perm, const uint32_t *__restrict__ off,
                                                    uint16_t *__restrict__ prog, uint8_t *__restrict__ qhdr,
                                                    uint32_t *__restrict__ status, const uint32_t *__restrict__ len4,
                                                    const uint16_t *__restrict__ stage, const uint8_t *__restrict__ hdr0) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    const uint32_t q = perm[i];
    const uint32_t o0 = off[i], o1 = off[i + 1];
    uint16_t *dst = prog + (size_t)o0 * 4;
    const int cap = (int)(o1 - o0) * 4;
    const uint32_t n = len4[q];
    if (n * 4u <= (uint32_t)SNAQ_STAGE_WORDS) {
        // short program: the classification pass kept it (NULL padded); a pad chunk, if any, is all NULL
        const uint2 *st = reinterpret_cast<const uint2 *>(stage + (size_t)q * SNAQ_STAGE_WORDS);
        uint2 *d2 = reinterpret_cast<uint2 *>(dst);
        const unsigned nul2 = (unsigned)T.n_xe * 0x10001u;
        for (uint32_t k = 0; k < o1 - o0; k++) d2[k] = k < n ? st[k] : make_uint2(nul2, nul2);
        qhdr[i] = hdr0[q];
        return;
    }
    uint16_t tx[4];
    load_taxa(taxa, q, tx);
    SnaqWriter w{dst, 0, cap};
    uint8_t hdr = 0;
    int e = snaq_build_quartet(T, tx, w, &hdr, nullptr);
    if (e) { atomicMax(status, (uint32_t)e); w.n = 0; hdr = 0; }
    else if (w.n > cap) { atomicMax(status, (uint32_t)SNAQ_B_OVERFLOW); w.n = cap; }
    for (int k = w.n; k < cap; k++) dst[k] = (uint16_t)T.n_xe;      // NULL slot (value 0.0)
    qhdr[i] = hdr;
}

// ---- program deduplication (opt-in: SNAQ_OPT_DEDUP) ----------------------------------------------------------------
// The expected CFs of a quartet depend on its program only (split, internal path, cycle terms), and the pseudo-deviance is
// linear in the weights 100*obsCF: quartets with the SAME program can be evaluated once with the SUM of their weights.
// After the sort by program hash, equal programs are neighbours: flag the first of every run (exact comparison of class,
// length and words), number the runs, copy one program per run into a second, compressed table, and add up the weights.
__global__ void __launch_bounds__(256) k_dedup_flags(const uint32_t *__restrict__ off, const uint2 *__restrict__ prog2,
                                                     const uint8_t *__restrict__ qhdr, int64_t nq, int64_t pad_at,
                                                     uint32_t *__restrict__ flags) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    uint32_t f = 1;
    if (i > 0) {
        const uint32_t a0 = off[i - 1], a1 = off[i], b1 = off[i + 1];
        const uint32_t la = a1 - a0 - (i - 1 == pad_at ? 1u : 0u), lb = b1 - a1 - (i == pad_at ? 1u : 0u);
        const unsigned ha = qhdr[i - 1], hb = qhdr[i];
        if (la == lb && SNAQ_HDR_KIND(ha) == SNAQ_HDR_KIND(hb) && SNAQ_HDR_TERMS(ha) == SNAQ_HDR_TERMS(hb)) {
            f = 0;
            for (uint32_t c = 0; c < la; c++) {
                const uint2 x = prog2[a0 + c], y = prog2[a1 + c];
                if (x.x != y.x || x.y != y.y) { f = 1; break; }
            }
        }
    }
    flags[i] = f;
}
// uidx = inclusive scan of flags (1-based run number).  Per run: first position, length (+1 pad chunk at pad_u), header.
__global__ void __launch_bounds__(256) k_dedup_heads(const uint32_t *__restrict__ flags, const uint32_t *__restrict__ uidx,
                                                     const uint32_t *__restrict__ off, const uint8_t *__restrict__ qhdr, int64_t nq,
                                                     int64_t pad_at, int64_t pad_u, uint32_t *__restrict__ rstart,
                                                     uint32_t *__restrict__ ulen, uint8_t *__restrict__ qhdr_u) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq || !flags[i]) return;
    const uint32_t u = uidx[i] - 1u;
    rstart[u] = (uint32_t)i;
    ulen[u] = off[i + 1] - off[i] - (i == pad_at ? 1u : 0u) + ((int64_t)u == pad_u ? 1u : 0u);
    qhdr_u[u] = qhdr[i];
}
__global__ void __launch_bounds__(256) k_dedup_copy(const uint32_t *__restrict__ rstart, const uint32_t *__restrict__ off,
                                                    const uint2 *__restrict__ prog2, const uint32_t *__restrict__ off_u, int64_t nu,
                                                    int64_t pad_u, unsigned nullslot, uint2 *__restrict__ prog_u) {
    const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= nu) return;
    const uint32_t src = off[rstart[u]], dst = off_u[u], n = off_u[u + 1] - dst - (u == pad_u ? 1u : 0u);
    for (uint32_t c = 0; c < n; c++) prog_u[dst + c] = prog2[src + c];
    if (u == pad_u) prog_u[dst + n] = make_uint2(nullslot * 0x10001u, nullslot * 0x10001u);
}
// error-free addition (run-summed weights: k_wu_combine)
__device__ __forceinline__ void two_sum(double a, double b, double &s, double &e) {
    s = __dadd_rn(a, b);
    const double bb = __dsub_rn(s, a);
    e = __dadd_rn(__dsub_rn(a, __dsub_rn(s, bb)), __dsub_rn(b, bb));
}
// W[i] (all classes) and, for the class-A quartets with one- and two-chunk programs [0,nA), the pair
// WA[i] = (W0, W1) that the single-call kernel streams; their constants W3 are summed once into wpart (one
// partial per block, fixed order)
// sampleCFfromCI! (src/bootstrap.jl:52-70) for every quartet: three uniforms from a counter-based generator keyed by the
// seed and the quartet's index in the data, c_i = (hi_i - lo_i) * U_i + lo_i, renormalised.  No fused multiply-add: the
// host restatement (numpy) gives the same bits.
__device__ __forceinline__ unsigned long long snaq_splitmix64(unsigned long long z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__global__ void __launch_bounds__(256) k_resample_ci(const double *__restrict__ lo, const double *__restrict__ hi, const int64_t nq,
                                                     const int64_t q_begin, const unsigned long long seed, double *__restrict__ obs) {
    const int64_t q = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (q >= nq) return;
    double c[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const unsigned long long ctr = (unsigned long long)(3 * (q_begin + q) + i);
        const double u = (double)(snaq_splitmix64(seed + ctr * 0x9E3779B97F4A7C15ull) >> 11) * 0x1.0p-53;
        c[i] = __dadd_rn(__dmul_rn(__dadd_rn(hi[3 * q + i], -lo[3 * q + i]), u), lo[3 * q + i]);
    }
    const double suma = __dadd_rn(__dadd_rn(c[0], c[1]), c[2]);
#pragma unroll
    for (int i = 0; i < 3; i++) obs[3 * q + i] = __ddiv_rn(c[i], suma);
}

__global__ void __launch_bounds__(256) k_weights(const uint32_t *__restrict__ perm, const uint8_t *__restrict__ qhdr,
                                                 const double *__restrict__ obs, const uint8_t *__restrict__ sampled,
                                                 int64_t nq, int64_t nA, int64_t nAB, double *__restrict__ W, double2 *__restrict__ WA,
                                                 double *__restrict__ wpart, const uint32_t *__restrict__ rec_off, uint32_t ntAB,
                                                 unsigned char *__restrict__ pack) {
    // WA[i] = (W0, W1) for every tree-like quartet i < nAB (k_eval_direct / k_eval_grad stream [0, nA), k_eval_runs all of
    // them); wpart[2 b] = the block's sum of W3 over the short additive programs [0, nA), wpart[2 b + 1] = over every
    // quartet that is not a bad-diamond-I 4-cycle (those keep their W3 with the per-quartet penalty rule)
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double c3 = 0.0, c3b = 0.0;
    if (i < nq) {
        const int64_t q = perm[i];
        double o[3] = {obs[3 * q], obs[3 * q + 1], obs[3 * q + 2]};
        double w[4];
        const unsigned h = qhdr[i];
        snaq_weights(h, o, sampled ? sampled[q] : 1, w);
        reinterpret_cast<double4 *>(W)[i] = make_double4(w[0], w[1], w[2], w[3]);
        if (i < nAB) WA[i] = make_double2(w[0], w[1]);
        if (pack) {                                   // the packed tile stream of k_eval_runs carries its own copy of the weights
            if (i < nAB) {
                const uint32_t t = (uint32_t)(i / RN_TQ), j = (uint32_t)(i % RN_TQ);
                reinterpret_cast<double2 *>(pack + (size_t)rec_off[t] * 16)[j] = make_double2(w[0], w[1]);
            } else {
                const uint32_t t = ntAB + (uint32_t)((i - nAB) / RN_TQC), j = (uint32_t)((i - nAB) % RN_TQC);
                reinterpret_cast<double4 *>(pack + (size_t)rec_off[t] * 16)[j] = make_double4(w[0], w[1], w[2], w[3]);
            }
        }
        if (i < nA) c3 = w[3];
        if (SNAQ_HDR_KIND(h) != SNAQ_KIND_T5BD1) c3b = w[3];
    }
    __shared__ double red[16];
    for (int o = 16; o > 0; o >>= 1) { c3 += __shfl_down_sync(0xffffffffu, c3, o); c3b += __shfl_down_sync(0xffffffffu, c3b, o); }
    if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5] = c3; red[8 + (threadIdx.x >> 5)] = c3b; }
    __syncthreads();
    if (threadIdx.x < 2) {
        const double *r = red + 8 * threadIdx.x;
        wpart[2 * blockIdx.x + threadIdx.x] = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    }
}

// parity descriptors: which / typeHyb per hybrid / formula, original order
__global__ void __launch_bounds__(128) k_describe(SnaqTables T, const uint16_t *__restrict__ taxa, int64_t nq,
                                                  int8_t *__restrict__ which, int8_t *__restrict__ types,
                                                  int nhyb, int8_t *__restrict__ formula) {
    int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    uint16_t tx[4];
    load_taxa(taxa, q, tx);
    SnaqWriter w{nullptr, 0, 0};
    SnaqQuartetInfo info;
    uint8_t hdr;
    snaq_build_quartet(T, tx, w, &hdr, &info);
    which[q] = info.which;
    for (int i = 0; i < 3; i++) formula[3 * q + i] = info.formula[i];
    for (int i = 0; i < nhyb; i++) types[q * nhyb + i] = info.type[i];
}

// hasEdge (updateHasEdge!, src/pseudolik.jl:392-464) and the "keeps a bad diamond I whole" flag that decides the form
// of indexht (parameters!(qnet,net), src/snaq_optimization.jl:120-146), original quartet order
__global__ void __launch_bounds__(128) k_hasedge(SnaqTables T, const uint16_t *__restrict__ taxa, int64_t nq, uint8_t *__restrict__ hasedge,
                                                 uint8_t *__restrict__ bd1, uint32_t *__restrict__ status) {
    int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    uint16_t tx[4];
    load_taxa(taxa, q, tx);
    uint32_t bits[SNAQ_HE_WORDS(SNAQ_MAX_SLOT + 1)];
    int flag = 0;
    const int e = snaq_hasedge_quartet(T, tx, bits, &flag);
    if (e) atomicMax(status, (uint32_t)e);
    for (int i = 0; i < T.nparams; i++) hasedge[q * T.nparams + i] = e ? 0 : (uint8_t)((bits[i >> 5] >> (i & 31)) & 1u);
    bd1[q] = (uint8_t)flag;
}

// the same for ANY level-1 network: every thread simulates extractQuartet! for its quartets on a private integer copy of the
// network (snaq_hesim.h); zpos: position in net.edge of the edge behind every gammaz slot (orders indexht)
__global__ void __launch_bounds__(64) k_hasedge_sim(SnaqSimNet net, const uint16_t *__restrict__ taxa, int64_t nq, uint8_t *__restrict__ hasedge,
                                                    uint8_t *__restrict__ bd1, int16_t *__restrict__ zpos, int nz, uint32_t *__restrict__ status) {
    SnaqSim g;
    uint32_t bits[SNAQ_HE_WORDS(SNAQ_MAX_SLOT + 1)];
    int16_t zp[64];
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += (int64_t)gridDim.x * blockDim.x) {
        uint16_t tx[4];
        load_taxa(taxa, q, tx);
        int flag = 0;
        const int e = snaq_hasedge_sim(net, tx, g, bits, &flag, nz > 0 ? zp : nullptr);
        if (e) atomicMax(status, (uint32_t)e);
        for (int i = 0; i < net.nparams; i++) hasedge[q * net.nparams + i] = e ? 0 : (uint8_t)((bits[i >> 5] >> (i & 31)) & 1u);
        for (int i = 0; i < nz; i++) zpos[q * nz + i] = e ? (int16_t)-1 : zp[i];
        bd1[q] = (uint8_t)flag;
    }
}

// q.qnet.edge of ONE quartet (edge numbers, in order): one thread runs the pruning
__global__ void k_quartet_edges(SnaqSimNet net, const uint16_t *__restrict__ taxa, int64_t q, int32_t *__restrict__ out, int cap,
                                int32_t *__restrict__ n) {
    SnaqSim g;
    uint16_t tx[4];
    load_taxa(taxa, q, tx);
    *n = snaq_quartet_edges_sim(net, tx, g, out, cap);
}

// --------------------------------------------------------------------------------------------
// evaluation
// --------------------------------------------------------------------------------------------
// Bounds-checked build (make libsnaqb200_dbg.so, -DSNAQ_BOUNDS): compute-sanitizer is not available on the GPU pool, so the
// debug library counts, on the device, every slot index outside the expanded parameter vector and every tile record /
// run / chunk index outside its stage (snaq_debug_bounds).  Any program word read from the wrong place is such an index.
#ifdef SNAQ_BOUNDS
__device__ unsigned long long g_bounds_bad = 0ull;
__device__ unsigned long long g_bounds_checks = 0ull;
__device__ int g_dbg_nfull = 0;
#define SNAQ_BCHK(cond) do { atomicAdd(&g_bounds_checks, 1ull); if (!(cond)) atomicAdd(&g_bounds_bad, 1ull); } while (0)
#else
#define SNAQ_BCHK(cond) do { } while (0)
#endif
struct XLane {   // candidate = lane; xs is [slot][32] (already offset by the lane)
    const double *xs;
    __device__ __forceinline__ double operator()(unsigned s) const {
#ifdef SNAQ_BOUNDS
        SNAQ_BCHK((int)s < g_dbg_nfull);
#endif
        return xs[s * 32u];
    }
};
struct XRow {    // one candidate's xe, contiguous
    const double *x;
    __device__ __forceinline__ double operator()(unsigned s) const {
#ifdef SNAQ_BOUNDS
        SNAQ_BCHK((int)s < g_dbg_nfull);
#endif
        return x[s];
    }
};

#ifdef SNAQ_BOUNDS
__device__ __forceinline__ unsigned dbg_slot(unsigned s) { SNAQ_BCHK((int)s < g_dbg_nfull); return s; }
#endif
__device__ __forceinline__ unsigned range_check(double v, double hi) { return (v >= 0.0 && v <= hi) ? 0u : SNAQ_S_RANGE; }

// Expanded parameter vectors (snaq_expand_row); also the one place where the bounds of the reference's
// update! are checked on the device (src/snaq_optimization.jl:213,221,228).
//   tiled == 0: XF[p][row]                       (thread-per-quartet kernel)
//   tiled == w (32 or 64): XF[p/w][row][p%w], P padded up to a multiple of w with copies of the last candidate:
//               the batched kernels copy one group's tile with a single TMA bulk copy
__global__ void __launch_bounds__(256) k_expand_x(SnaqTables T, const double *__restrict__ X, const double *__restrict__ xconst,
                                                  int P, int nh, int nt, int tiled, double *__restrict__ XF,
                                                  uint32_t *__restrict__ status) {
    grid_launch_dependents();
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const SnaqMath M{g_math_tab, g_math_tab + 16, g_math_tab + 32};
    if (!tiled) {
        if (idx >= (int64_t)P * T.n_full) return;
        const int p = (int)(idx / T.n_full), row = (int)(idx - (int64_t)p * T.n_full);
        const double *x = X + (size_t)p * T.nparams;
        if (row < T.nparams) {
            const unsigned st = range_check(x[row], (row >= nh && row < nh + nt) ? 1.0e300 : 1.0);
            if (st) atomicOr(status, st);
        }
        XF[idx] = snaq_expand_row(T, M, x, xconst, row);
    } else {
        const int Ppad = (P + tiled - 1) / tiled * tiled;
        if (idx >= (int64_t)Ppad * T.n_full) return;
        const int lane = (int)(idx % tiled);
        const int64_t gr = idx / tiled;                    // group * n_full + row
        const int g = (int)(gr / T.n_full), row = (int)(gr - (int64_t)g * T.n_full);
        const int p = min(g * tiled + lane, P - 1);
        const double *x = X + (size_t)p * T.nparams;
        if (row < T.nparams) {
            const unsigned st = range_check(x[row], (row >= nh && row < nh + nt) ? 1.0e300 : 1.0);
            if (st) atomicOr(status, st);
        }
        XF[idx] = snaq_expand_row(T, M, x, xconst, row);
    }
}

struct ClassRanges {       // sorted quartet positions [0,nA) class A, [nA,nA+nB) class B, rest class C
    int64_t nA1;           // the first nA1 class-A quartets have one-chunk programs: off[i] == i
    int64_t nA, nB, nC;
    int chA, chB, chC;     // chunks (blocks) per class
    int64_t nA2;           // the next nA2 class-A quartets have two-chunk programs at chunk a2base + 2 j
    uint32_t a2base;
    int noreuse;           // 1: do not reuse t1 / exp / log / terms between neighbouring quartets (measurement aid)
    const double *pen;     // deduplicated table: per-entry penalty data {E0, E1, E2, n}, else nullptr
};

/*
 * grid (nchunks, ngroups); block = WARPS*32 threads.  Block (cb, g) evaluates one chunk of quartets
 * of ONE class for candidates [32g, 32g+32).  Dynamic shared memory (128-byte aligned):
 *   mt   [48] f64          exp / log tables (three 128-byte rows: conflict-free lookups)
 *   bar  mbarrier
 *   xs   [n_full][32] f64  the group's expanded parameter vectors      <- one TMA bulk copy
 *   sW   [chunk][4] f64    weights of the chunk's quartets             <- one TMA bulk copy
 *   sprog 8-byte chunks    programs of the chunk's quartets            <- one TMA bulk copy
 *   red  [WARPS][32] f64
 *   soff [chunk+1] u32     program offsets relative to the chunk (4-word units)
 */
struct LanesSmem {
    size_t mt, bar, xs, sW, sprog, red, soff, total;
};
__host__ __device__ inline LanesSmem lanes_smem_layout(int n_full, int chunk, int warps, uint32_t maxlen) {
    LanesSmem L;
    size_t o = 0;
    L.mt = o; o += SNAQ_MATH_DOUBLES * 8;            // 384
    L.bar = o; o += 128;
    L.xs = o; o += (size_t)n_full * 256;
    L.sW = o; o += (size_t)chunk * 32;
    L.sprog = o; o += ((size_t)chunk * maxlen + 2) * 8;
    o = (o + 15) & ~size_t(15);
    L.red = o; o += (size_t)warps * 256;
    L.soff = o; o += ((size_t)chunk + 1) * 4;
    L.total = (o + 15) & ~size_t(15);
    return L;
}

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) k_eval_lanes(const uint32_t *__restrict__ off, const uint16_t *__restrict__ prog,
                                                         const uint8_t *__restrict__ qhdr, const double *__restrict__ W,
                                                         ClassRanges cr, int chunk, uint32_t maxlen, const double *__restrict__ XFT,
                                                         int n_full, double *__restrict__ partial, int Ppad,
                                                         uint32_t *__restrict__ status) {
    extern __shared__ __align__(128) unsigned char smem[];
    const LanesSmem L = lanes_smem_layout(n_full, chunk, WARPS, maxlen);
    double *mt = reinterpret_cast<double *>(smem + L.mt);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + L.bar);
    double *xs = reinterpret_cast<double *>(smem + L.xs);
    double *sW = reinterpret_cast<double *>(smem + L.sW);
    uint2 *sprog_raw = reinterpret_cast<uint2 *>(smem + L.sprog);
    double *red = reinterpret_cast<double *>(smem + L.red);
    uint32_t *soff = reinterpret_cast<uint32_t *>(smem + L.soff);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = blockIdx.y;
    grid_launch_dependents();
    const bool noreuse = cr.noreuse != 0;          // measurement aid (SNAQ_REUSE=0): every quartet evaluated in full
    // heavy classes first: C, B, then A
    int cls, cb = blockIdx.x;
    int64_t qbase, qend;
    if (cb < cr.chC) { cls = 2; qbase = cr.nA + cr.nB; qend = qbase + cr.nC; }
    else if (cb < cr.chC + cr.chB) { cls = 1; cb -= cr.chC; qbase = cr.nA; qend = qbase + cr.nB; }
    else { cls = 0; cb -= cr.chC + cr.chB; qbase = 0; qend = cr.nA; }
    const int64_t q0 = qbase + (int64_t)cb * chunk;
    const int nqc = (int)min((int64_t)chunk, qend - q0);
    const uint32_t base = off[q0];
    const uint32_t nchunks = off[q0 + nqc] - base;
    // programs start at an 8-byte boundary; the bulk copy needs 16: start one chunk earlier if odd
    const uint32_t base16 = base & ~1u;
    const unsigned prog_bytes = (((base - base16) + nchunks) * 8u + 15u) & ~15u;
    if (tid == 0) mbar_init(bar, 1);
    __syncthreads();
    if (tid == 0) {
        const unsigned xs_bytes = (unsigned)n_full * 256u, w_bytes = (unsigned)nqc * 32u;
        mbar_expect_tx(bar, xs_bytes + w_bytes + prog_bytes);
        tma_load_1d(sW, W + (size_t)q0 * 4, w_bytes, bar);
        tma_load_1d(sprog_raw, reinterpret_cast<const uint2 *>(prog) + base16, prog_bytes, bar);
        grid_dependency_wait();                                  // the tile is k_expand_x's output (same step)
        tma_load_1d(xs, XFT + (size_t)g * n_full * 32, xs_bytes, bar);
    }
    const SnaqMath M = load_math_tables(mt, tid);
    for (int i = tid; i <= nqc; i += WARPS * 32) soff[i] = off[q0 + i] - base;
    const uint2 *sprog = sprog_raw + (base - base16);
    unsigned st = 0;
    mbar_wait(bar, 0);
    __syncthreads();
    double acc = 0.0;
    const double *xl = xs + lane;
    // One PRMT per slot builds the byte offset slot*256 + lane*8 of xf[slot] for this lane.
    const unsigned lane8 = (unsigned)lane * 8u;
    const char *xb = reinterpret_cast<const char *>(xs);
#ifdef SNAQ_BOUNDS
#define XLO(w) (*(xl + 32u * dbg_slot((w) & 0xffffu)))
#define XHI(w) (*(xl + 32u * dbg_slot((w) >> 16)))
#else
#define XLO(w) (*reinterpret_cast<const double *>(xb + __byte_perm((w), lane8, 0x5104)))
#define XHI(w) (*reinterpret_cast<const double *>(xb + __byte_perm((w), lane8, 0x5324)))
#endif
    // a warp takes a CONTIGUOUS share of the chunk: neighbours in the sorted table share programs / terms
    const int wper = (nqc + WARPS - 1) / WARPS, wbeg = min(warp * wper, nqc), wend = min(wbeg + wper, nqc);
    if (cls == 0) {
        // class A: t1 = min(10, sum of signed pairs).  The class is sorted by program, and a quartet's expected CFs depend
        // on its program only: while the program repeats (quartets that differ in taxa of the same clades) t1, exp and log
        // are reused and the quartet costs its two weight FMAs (the program is warp-uniform: the test does not diverge).
        auto pair_sum = [&](int i) {
            uint32_t c = soff[i];
            const uint32_t o1 = soff[i + 1];
            uint2 w = sprog[c];
            double t0 = XLO(w.x) - XHI(w.x), t1 = XLO(w.y) - XHI(w.y);
#pragma unroll 1
            for (++c; c < o1; ++c) {
                w = sprog[c];
                t0 += XLO(w.x) - XHI(w.x);
                t1 += XLO(w.y) - XHI(w.y);
            }
            return snaq_clamp10(t0 + t1);
        };
        uint2 mw = make_uint2(0xffffffffu, 0xffffffffu), mw2 = mw;
        double m_t = 0.0, m_l = 0.0;
        int i = wbeg;
        if (q0 + nqc <= cr.nA1) {
            // every program of this chunk is ONE chunk at sprog[i]: no offsets, two quartets per iteration (the loads of the
            // second do not wait for the first)
            auto one = [&](const uint2 &w, const double4 &wa) {
                if (noreuse || w.x != mw.x || w.y != mw.y) {
                    m_t = snaq_clamp10((XLO(w.x) - XHI(w.x)) + (XLO(w.y) - XHI(w.y)));
                    m_l = snaq_log_pos(fma(SNAQ_K_M23, snaq_exp(-m_t, M), 1.0), M);
                    mw = w;
                }
                acc += fma(wa.x, m_l, -fma(wa.y, m_t + SNAQ_K_LOG3, wa.w));
            };
#pragma unroll 1
            for (; i + 1 < wend; i += 2) {
                const uint2 wA = sprog[i], wB = sprog[i + 1];
                const double4 waA = reinterpret_cast<const double4 *>(sW)[i], waB = reinterpret_cast<const double4 *>(sW)[i + 1];
                one(wA, waA);
                one(wB, waB);
            }
            if (i < wend) { one(sprog[i], reinterpret_cast<const double4 *>(sW)[i]); i = wend; }
            mw2 = make_uint2(0xfffffffeu, 1u);                     // consistent with the general loop below (not entered)
        } else if (q0 >= cr.nA1 && q0 + nqc <= cr.nA1 + cr.nA2) {
            // every program of this chunk is TWO chunks at sprog[2 i]
            auto two = [&](const uint2 &w, const uint2 &w2, const double4 &wa) {
                if (noreuse || w.x != mw.x || w.y != mw.y || w2.x != mw2.x || w2.y != mw2.y) {
                    const double t0 = (XLO(w.x) - XHI(w.x)) + (XLO(w2.x) - XHI(w2.x)), t1 = (XLO(w.y) - XHI(w.y)) + (XLO(w2.y) - XHI(w2.y));
                    m_t = snaq_clamp10(t0 + t1);
                    m_l = snaq_log_pos(fma(SNAQ_K_M23, snaq_exp(-m_t, M), 1.0), M);
                    mw = w; mw2 = w2;
                }
                acc += fma(wa.x, m_l, -fma(wa.y, m_t + SNAQ_K_LOG3, wa.w));
            };
#pragma unroll 1
            for (; i + 1 < wend; i += 2) {
                const uint2 a0 = sprog[2 * i], a1 = sprog[2 * i + 1], b0 = sprog[2 * i + 2], b1 = sprog[2 * i + 3];
                const double4 waA = reinterpret_cast<const double4 *>(sW)[i], waB = reinterpret_cast<const double4 *>(sW)[i + 1];
                two(a0, a1, waA);
                two(b0, b1, waB);
            }
            if (i < wend) { two(sprog[2 * i], sprog[2 * i + 1], reinterpret_cast<const double4 *>(sW)[i]); i = wend; }
        }
#pragma unroll 1
        for (; i < wend; i++) {
            const uint32_t c0 = soff[i];
            const uint32_t len = soff[i + 1] - c0;
            const uint2 w = sprog[c0];
            const uint2 w2 = (len == 2u) ? sprog[c0 + 1] : make_uint2(0xfffffffeu, len);
            if (noreuse || len > 2u || w.x != mw.x || w.y != mw.y || w2.x != mw2.x || w2.y != mw2.y) {
                if (len == 1u) m_t = snaq_clamp10((XLO(w.x) - XHI(w.x)) + (XLO(w.y) - XHI(w.y)));
                else m_t = pair_sum(i);
                const double y = snaq_exp(-m_t, M);
                m_l = snaq_log_pos(fma(SNAQ_K_M23, y, 1.0), M);
                mw = w; mw2 = w2;
                if (len > 2u) mw = make_uint2(0xffffffffu, 0xffffffffu);          // longer programs are not compared: never reused
            }
            const double4 wa = reinterpret_cast<const double4 *>(sW)[i];
            acc += fma(wa.x, m_l, -fma(wa.y, m_t + SNAQ_K_LOG3, wa.w));
        }
    } else if (cls == 1) {
        // class B: tree-like with type-2 / type-4 terms (same arithmetic as snaq_tree_terms_t1).  The class is sorted by
        // term, so the next quartet usually has the SAME term (cycle and hit positions): the value just computed is reused
        // (the program is warp-uniform: the test does not diverge).
        unsigned m4_k0 = 0xffffffffu, m4_k1 = 0xffffffffu, m2_k0 = 0xffffffffu, m2_k1 = 0xffffffffu;
        double m4_l4 = 0.0, m2_v = 0.0;
        unsigned mB_len = 0xffffffffu;
        uint2 mB_w0 = make_uint2(0u, 0u), mB_w1 = mB_w0, mB_w2 = mB_w0;
        double mB_t = 0.0, mB_lm = 0.0;
#pragma unroll 1
        for (int i = wbeg; i < wend; i++) {
            // the WHOLE program is often the previous one (the class is sorted by term, then by program): reuse t and log(major)
            const uint32_t bc0 = soff[i], blen = soff[i + 1] - bc0;
            const double4 w4 = reinterpret_cast<const double4 *>(sW)[i];
            bool same = (blen == mB_len) && !noreuse;
            if (same) {
                const uint2 a0 = sprog[bc0], a1 = blen > 1u ? sprog[bc0 + 1] : mB_w1, a2 = blen > 2u ? sprog[bc0 + 2] : mB_w2;
                same = a0.x == mB_w0.x && a0.y == mB_w0.y && a1.x == mB_w1.x && a1.y == mB_w1.y && a2.x == mB_w2.x && a2.y == mB_w2.y;
            }
            if (same) {
                acc += fma(w4.x, (w4.x != 0.0) ? mB_lm : 0.0, -fma(w4.y, mB_t + SNAQ_K_LOG3, w4.w));
                continue;
            }
            const uint16_t *pc = reinterpret_cast<const uint16_t *>(sprog + bc0);
            const unsigned cw = pc[0];
            const int npairs = (int)(cw & 255u), nterms = (int)(cw >> 8);
            double R = 0.0;
            int k = 1;
#pragma unroll 1
            for (int j = 0; j < npairs; j++, k += 2) {
                const unsigned wv = (unsigned)pc[k] | ((unsigned)pc[k + 1] << 16);
                R += XLO(wv) - XHI(wv);
            }
            double Bn = 0.0, Bf = 0.0;
#pragma unroll 1
            for (int it = 0; it < nterms; it++) {
                const unsigned th = pc[k];
                const double gq = XLO(th >> 4);
                if ((th & 3u) == SNAQ_T2) {
                    const unsigned wv = (unsigned)pc[k + 1] | ((unsigned)pc[k + 2] << 16);
                    k += 3;
                    if (noreuse || th != m2_k0 || wv != m2_k1) {
                        const double e = snaq_clamp10(XLO(wv) - XHI(wv));
                        const double gh = (th & 4u) ? 1.0 - gq : gq;
                        m2_v = snaq_clamp10(-snaq_log(1.0 - gh * (1.0 - snaq_exp(-e, M)), M));
                        m2_k0 = th; m2_k1 = wv;
                    }
                    R += m2_v;
                } else {
                    const unsigned wv = (unsigned)pc[k + 1] | ((unsigned)pc[k + 2] << 16);
                    const unsigned kk0 = (th & ~4u) | ((unsigned)pc[k + 3] << 16);
                    const int nchain = pc[k + 4];
                    k += 5;
                    double ch = 0.0;
#pragma unroll 1
                    for (int j = 0; j < nchain; j++, k += 2) {
                        const unsigned wc = (unsigned)pc[k] | ((unsigned)pc[k + 1] << 16);
                        ch += XLO(wc) - XHI(wc);
                    }
                    ch = snaq_clamp10(ch);
                    if (noreuse || kk0 != m4_k0 || wv != m4_k1) {
                        const double clo = XLO(wv), chi = XHI(wv), ck = XLO((unsigned)pc[k - 2 * nchain - 2]);
                        const double ya = snaq_exp(-snaq_clamp10(clo), M), yb = snaq_exp(-snaq_clamp10(ck - chi), M);
                        const double ye = snaq_exp(-snaq_clamp10(chi - clo), M);
                        const double ga = 1.0 - gq, gb = gq;
                        m4_l4 = -snaq_log(gb * gb * yb + ga * gb * (3.0 - ye) + ga * ga * ya, M);
                        m4_k0 = kk0; m4_k1 = wv;
                    }
                    double l4 = m4_l4;
                    if (l4 < SNAQ_MINLEN) st |= SNAQ_S_NEGLEN;
                    l4 = snaq_clamp10(l4);
                    const double B = snaq_clamp10(ch + l4);
                    if (B < SNAQ_MINLEN) st |= SNAQ_S_NEGLEN;
                    if (th & 4u) Bf = B; else Bn = B;
                }
            }
            double t = snaq_clamp10(Bn + R);
            if (t < SNAQ_MINLEN) st |= SNAQ_S_NEGLEN;
            t = snaq_clamp10(t + Bf);
            if (t < SNAQ_MINLEN) st |= SNAQ_S_NEGLEN;
            const double y = snaq_exp(-t, M);
            const double major = fma(SNAQ_K_M23, y, 1.0);
            if (major < 0.0) st |= SNAQ_S_NEGLEN;
            mB_lm = snaq_log(major, M);
            mB_t = t;
            if (blen <= 3u) {                                            // longer programs are not remembered
                mB_len = blen;
                mB_w0 = sprog[bc0];
                mB_w1 = blen > 1u ? sprog[bc0 + 1] : make_uint2(0u, 0u);
                mB_w2 = blen > 2u ? sprog[bc0 + 2] : make_uint2(0u, 0u);
            } else mB_len = 0xffffffffu;
            acc += fma(w4.x, (w4.x != 0.0) ? mB_lm : 0.0, -fma(w4.y, t + SNAQ_K_LOG3, w4.w));      // obs == 0 contributes 0 (src/pseudolik.jl:1398)
        }
    } else {
        // class C: 4-cycle quartets.  T5: [gslot, C_a, C_m, C_b] in one chunk; all three CFs are positive.  Sorted by these
        // four slots: the three logs are reused while they repeat.
        const XLane xf{xl};
        unsigned m5_k0 = 0xffffffffu, m5_k1 = 0xffffffffu;
        double m5_l0 = 0.0, m5_l1 = 0.0, m5_l2 = 0.0;
        const uint8_t *hd = qhdr + q0;
#pragma unroll 1
        for (int i = wbeg; i < wend; i++) {
            const uint2 w = sprog[soff[i]];
            const double4 w4 = reinterpret_cast<const double4 *>(sW)[i];
            if (SNAQ_HDR_KIND(hd[i]) == SNAQ_KIND_T5) {
                if (noreuse || w.x != m5_k0 || w.y != m5_k1) {
                    const double g2 = XLO(w.x), ca = XHI(w.x), cm = XLO(w.y), cb = XHI(w.y);
                    const double g1 = 1.0 - g2;
                    const double y5 = snaq_exp(-snaq_clamp10(cm - ca), M), y6 = snaq_exp(-snaq_clamp10(cb - cm), M);
                    const double g13 = snaq_div3(g1), g23 = snaq_div3(g2);          // = g/3.0 bit for bit, no division
                    const double cf0 = g1 * (1.0 - 2.0 / 3.0 * y5) + g23 * y6;
                    const double cf1 = g13 * y5 + g2 * (1.0 - 2.0 / 3.0 * y6);
                    const double cf2 = g13 * y5 + g23 * y6;
                    m5_l0 = snaq_log_pos(cf0, M); m5_l1 = snaq_log_pos(cf1, M); m5_l2 = snaq_log_pos(cf2, M);
                    m5_k0 = w.x; m5_k1 = w.y;
                }
                double v = -w4.w;
                v = fma(w4.x, m5_l0, v);
                v = fma(w4.y, m5_l1, v);
                v = fma(w4.z, m5_l2, v);
                acc += v;
            } else {
                const uint16_t *pc = reinterpret_cast<const uint16_t *>(sprog + soff[i]);
                double cf[3];
                snaq_t5_cf(SNAQ_KIND_T5BD1, pc, xf, M, cf);
                acc += snaq_loglik_t5(cf, sW + 4 * i, M, st, cr.pen ? cr.pen + 4 * (q0 + i) : nullptr);
            }
        }
    }
#undef XLO
#undef XHI
    red[warp * 32 + lane] = acc;
    __syncthreads();
    if (warp == 0) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < WARPS; w++) s += red[w * 32 + lane];
        partial[(size_t)blockIdx.x * Ppad + 32 * g + lane] = s;
    }
    if (st) atomicOr(status, st);
}

/*
 * Persistent variant of k_eval_lanes: grid (gx, ngroups).  A block keeps its group's xf tile in shared memory and walks
 * the quartet chunks cb = blockIdx.x, blockIdx.x + gx, ...: the weights and programs of chunk k+1 arrive by TMA bulk
 * copies into the other of two stages while chunk k is evaluated.  The tile is loaded once per block instead of once
 * per chunk (it is 150 KB at 200 taxa), the per-block prologue and epilogue are paid once, and the second pass adds gx
 * rows instead of one row per chunk.  Chunks inside the one- and two-chunk class-A regions need no offset table.
 */
struct LanesPSmem {
    size_t mt, bar, xs, stage, stage_bytes, sprog_off, red, soff, total;
};
__host__ __device__ inline LanesPSmem lanesp_smem_layout(int n_full, int chunk, int warps, uint32_t maxlen) {
    LanesPSmem L;
    size_t o = 0;
    L.mt = o; o += SNAQ_MATH_DOUBLES * 8;
    L.bar = o; o += 128;
    L.xs = o; o += (size_t)n_full * 256;
    L.sprog_off = (size_t)chunk * 32;
    L.stage_bytes = (L.sprog_off + ((size_t)chunk * maxlen + 2) * 8 + 15) & ~size_t(15);
    L.stage = o; o += 2 * L.stage_bytes;
    L.red = o; o += (size_t)warps * 256;
    L.soff = o; o += ((size_t)chunk + 1) * 4;
    L.total = (o + 15) & ~size_t(15);
    return L;
}

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) k_eval_lanes_p(const uint32_t *__restrict__ off, const uint16_t *__restrict__ prog,
                                                           const uint8_t *__restrict__ qhdr, const double *__restrict__ W,
                                                           ClassRanges cr, int chunk, uint32_t maxlen, const double *__restrict__ XFT,
                                                           int n_full, double *__restrict__ partial, int Ppad,
                                                           uint32_t *__restrict__ status, const int LANESP_SUPER) {
    extern __shared__ __align__(128) unsigned char smem[];
    const LanesPSmem L = lanesp_smem_layout(n_full, chunk, WARPS, maxlen);
    double *mt = reinterpret_cast<double *>(smem + L.mt);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + L.bar);          // [0] xf tile, [1], [2] the two chunk stages
    double *xs = reinterpret_cast<double *>(smem + L.xs);
    double *red = reinterpret_cast<double *>(smem + L.red);
    uint32_t *soff = reinterpret_cast<uint32_t *>(smem + L.soff);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = blockIdx.y;
    const bool noreuse = cr.noreuse != 0;          // measurement aid (SNAQ_REUSE=0): every quartet evaluated in full
    const int nchunks_tot = cr.chA + cr.chB + cr.chC;
    const uint2 *prog2 = reinterpret_cast<const uint2 *>(prog);
    struct Meta { int cls; int64_t q0; int nqc; bool implicit; uint32_t mult, base, nch; };
    // heavy classes first: C, B, then A
    auto meta = [&](int cb) {
        Meta m;
        int64_t qbase, qend;
        if (cb < cr.chC) { m.cls = 2; qbase = cr.nA + cr.nB; qend = qbase + cr.nC; }
        else if (cb < cr.chC + cr.chB) { m.cls = 1; cb -= cr.chC; qbase = cr.nA; qend = qbase + cr.nB; }
        else { m.cls = 0; cb -= cr.chC + cr.chB; qbase = 0; qend = cr.nA; }
        m.q0 = qbase + (int64_t)cb * chunk;
        m.nqc = (int)min((int64_t)chunk, qend - m.q0);
        m.implicit = false; m.mult = 1u;
        if (m.cls == 0 && m.q0 + m.nqc <= cr.nA1) { m.implicit = true; m.base = (uint32_t)m.q0; m.nch = (uint32_t)m.nqc; }
        else if (m.cls == 0 && m.q0 >= cr.nA1 && m.q0 + m.nqc <= cr.nA1 + cr.nA2) {
            m.implicit = true; m.mult = 2u; m.base = cr.a2base + 2u * (uint32_t)(m.q0 - cr.nA1); m.nch = 2u * (uint32_t)m.nqc;
        } else { m.base = off[m.q0]; m.nch = off[m.q0 + m.nqc] - m.base; }
        return m;
    };
    auto issue = [&](const Meta &m, int s) {                             // thread 0 only
        unsigned char *st = smem + L.stage + (size_t)s * L.stage_bytes;
        const uint32_t base16 = m.base & ~1u;                            // programs start at 8 bytes; the bulk copy needs 16
        const unsigned prog_bytes = (((m.base - base16) + m.nch) * 8u + 15u) & ~15u, w_bytes = (unsigned)m.nqc * 32u;
        mbar_expect_tx(&bar[1 + s], w_bytes + prog_bytes);
        tma_load_1d(st, W + (size_t)m.q0 * 4, w_bytes, &bar[1 + s]);
        tma_load_1d(st + L.sprog_off, prog2 + base16, prog_bytes, &bar[1 + s]);
    };
    if (tid == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); mbar_init(&bar[2], 1); }
    __syncthreads();
    grid_launch_dependents();
    if (tid == 0) {
        if ((int)blockIdx.x * LANESP_SUPER < nchunks_tot) issue(meta((int)blockIdx.x * LANESP_SUPER), 0);
        mbar_expect_tx(&bar[0], (unsigned)n_full * 256u);
        grid_dependency_wait();                                  // the tile is k_expand_x's output (same step)
        tma_load_1d(xs, XFT + (size_t)g * n_full * 32, (unsigned)n_full * 256u, &bar[0]);
    }
    const SnaqMath M = load_math_tables(mt, tid);
    unsigned st = 0;
    mbar_wait(&bar[0], 0);
    __syncthreads();
    double acc = 0.0;
    const double *xl = xs + lane;
    // One PRMT per slot builds the byte offset slot*256 + lane*8 of xf[slot] for this lane.
    const unsigned lane8 = (unsigned)lane * 8u;
    const char *xb = reinterpret_cast<const char *>(xs);
#ifdef SNAQ_BOUNDS
#define XLO(w) (*(xl + 32u * dbg_slot((w) & 0xffffu)))
#define XHI(w) (*(xl + 32u * dbg_slot((w) >> 16)))
#else
#define XLO(w) (*reinterpret_cast<const double *>(xb + __byte_perm((w), lane8, 0x5104)))
#define XHI(w) (*reinterpret_cast<const double *>(xb + __byte_perm((w), lane8, 0x5324)))
#endif
    // A block walks SUPER-CHUNKS (LANESP_SUPER consecutive chunks; the host derives it from the table size only) sc = blockIdx.x, blockIdx.x + gx, ... and writes one
    // partial row per super-chunk: the sums do not depend on how many blocks the launch has (hence not on P).
    auto next_chunk = [&](int cb) {
        if ((cb % LANESP_SUPER) != LANESP_SUPER - 1 && cb + 1 < nchunks_tot) return cb + 1;
        const int nx = (cb / LANESP_SUPER + (int)gridDim.x) * LANESP_SUPER;
        return nx < nchunks_tot ? nx : -1;
    };
    int it = 0;
#pragma unroll 1
    for (int cbi = (int)blockIdx.x * LANESP_SUPER; cbi >= 0 && cbi < nchunks_tot; it++) {
        const int sidx = it & 1;
        const Meta m = meta(cbi);
        const int cnext = next_chunk(cbi);
        // the other stage was last read in the previous iteration, which ended with a __syncthreads: refill it
        if (tid == 0 && cnext >= 0) issue(meta(cnext), sidx ^ 1);
        const int cls = m.cls, nqc = m.nqc;
        const bool implicit = m.implicit;
        const uint32_t mult = m.mult;
        if (!implicit) for (int i = tid; i <= nqc; i += WARPS * 32) soff[i] = off[m.q0 + i] - m.base;
        auto SO = [&](int i) { return implicit ? (uint32_t)i * mult : soff[i]; };
        const unsigned char *stg = smem + L.stage + (size_t)sidx * L.stage_bytes;
        const double *sW = reinterpret_cast<const double *>(stg);
        const uint2 *sprog = reinterpret_cast<const uint2 *>(stg + L.sprog_off) + (m.base - (m.base & ~1u));
        const int64_t q0 = m.q0;
        mbar_wait(&bar[1 + sidx], (unsigned)(it >> 1) & 1u);
        __syncthreads();
        // a warp takes a CONTIGUOUS share of the chunk: neighbours in the sorted table share programs / terms
        const int wper = (nqc + WARPS - 1) / WARPS, wbeg = min(warp * wper, nqc), wend = min(wbeg + wper, nqc);
        if (cls == 0) {
            // class A: t1 = min(10, sum of signed pairs).  The class is sorted by program, and a quartet's expected CFs depend
            // on its program only: while the program repeats (quartets that differ in taxa of the same clades) t1, exp and log
            // are reused and the quartet costs its two weight FMAs (the program is warp-uniform: the test does not diverge).
            auto pair_sum = [&](int i) {
                uint32_t c = SO(i);
                const uint32_t o1 = SO(i + 1);
                uint2 w = sprog[c];
                double t0 = XLO(w.x) - XHI(w.x), t1 = XLO(w.y) - XHI(w.y);
#pragma unroll 1
                for (++c; c < o1; ++c) {
                    w = sprog[c];
                    t0 += XLO(w.x) - XHI(w.x);
                    t1 += XLO(w.y) - XHI(w.y);
                }
                return snaq_clamp10(t0 + t1);
            };
            uint2 mw = make_uint2(0xffffffffu, 0xffffffffu), mw2 = mw;
            double m_t = 0.0, m_l = 0.0;
            int i = wbeg;
            if (implicit && mult == 1u) {
                // every program of this chunk is ONE chunk at sprog[i]: no offsets, two quartets per iteration (the loads of the
                // second do not wait for the first)
                auto one = [&](const uint2 &w, const double4 &wa) {
                    if (noreuse || w.x != mw.x || w.y != mw.y) {
                        m_t = snaq_clamp10((XLO(w.x) - XHI(w.x)) + (XLO(w.y) - XHI(w.y)));
                        m_l = snaq_log_pos(fma(SNAQ_K_M23, snaq_exp(-m_t, M), 1.0), M);
                        mw = w;
                    }
                    acc += fma(wa.x, m_l, -fma(wa.y, m_t + SNAQ_K_LOG3, wa.w));
                };
#pragma unroll 1
                for (; i + 1 < wend; i += 2) {
                    const uint2 wA = sprog[i], wB = sprog[i + 1];
                    const double4 waA = reinterpret_cast<const double4 *>(sW)[i], waB = reinterpret_cast<const double4 *>(sW)[i + 1];
                    one(wA, waA);
                    one(wB, waB);
                }
                if (i < wend) { one(sprog[i], reinterpret_cast<const double4 *>(sW)[i]); i = wend; }
                mw2 = make_uint2(0xfffffffeu, 1u);                     // consistent with the general loop below (not entered)
            } else if (implicit && mult == 2u) {
                // every program of this chunk is TWO chunks at sprog[2 i]
                auto two = [&](const uint2 &w, const uint2 &w2, const double4 &wa) {
                    if (noreuse || w.x != mw.x || w.y != mw.y || w2.x != mw2.x || w2.y != mw2.y) {
                        const double t0 = (XLO(w.x) - XHI(w.x)) + (XLO(w2.x) - XHI(w2.x)), t1 = (XLO(w.y) - XHI(w.y)) + (XLO(w2.y) - XHI(w2.y));
                        m_t = snaq_clamp10(t0 + t1);
                        m_l = snaq_log_pos(fma(SNAQ_K_M23, snaq_exp(-m_t, M), 1.0), M);
                        mw = w; mw2 = w2;
                    }
                    acc += fma(wa.x, m_l, -fma(wa.y, m_t + SNAQ_K_LOG3, wa.w));
                };
#pragma unroll 1
                for (; i + 1 < wend; i += 2) {
                    const uint2 a0 = sprog[2 * i], a1 = sprog[2 * i + 1], b0 = sprog[2 * i + 2], b1 = sprog[2 * i + 3];
                    const double4 waA = reinterpret_cast<const double4 *>(sW)[i], waB = reinterpret_cast<const double4 *>(sW)[i + 1];
                    two(a0, a1, waA);
                    two(b0, b1, waB);
                }
                if (i < wend) { two(sprog[2 * i], sprog[2 * i + 1], reinterpret_cast<const double4 *>(sW)[i]); i = wend; }
            }
#pragma unroll 1
            for (; i < wend; i++) {
                const uint32_t c0 = SO(i);
                const uint32_t len = SO(i + 1) - c0;
                const uint2 w = sprog[c0];
                const uint2 w2 = (len == 2u) ? sprog[c0 + 1] : make_uint2(0xfffffffeu, len);
                if (noreuse || len > 2u || w.x != mw.x || w.y != mw.y || w2.x != mw2.x || w2.y != mw2.y) {
                    if (len == 1u) m_t = snaq_clamp10((XLO(w.x) - XHI(w.x)) + (XLO(w.y) - XHI(w.y)));
                    else m_t = pair_sum(i);
                    const double y = snaq_exp(-m_t, M);
                    m_l = snaq_log_pos(fma(SNAQ_K_M23, y, 1.0), M);
                    mw = w; mw2 = w2;
                    if (len > 2u) mw = make_uint2(0xffffffffu, 0xffffffffu);          // longer programs are not compared: never reused
                }
                const double4 wa = reinterpret_cast<const double4 *>(sW)[i];
                acc += fma(wa.x, m_l, -fma(wa.y, m_t + SNAQ_K_LOG3, wa.w));
            }
        } else if (cls == 1) {
            // class B: tree-like with type-2 / type-4 terms (same arithmetic as snaq_tree_terms_t1); terms are reused while
            // they repeat (see k_eval_lanes)
            unsigned m4_k0 = 0xffffffffu, m4_k1 = 0xffffffffu, m2_k0 = 0xffffffffu, m2_k1 = 0xffffffffu;
            double m4_l4 = 0.0, m2_v = 0.0;
            unsigned mB_len = 0xffffffffu;
            uint2 mB_w0 = make_uint2(0u, 0u), mB_w1 = mB_w0, mB_w2 = mB_w0;
            double mB_t = 0.0, mB_lm = 0.0;
#pragma unroll 1
            for (int i = wbeg; i < wend; i++) {
                // the WHOLE program is often the previous one (the class is sorted by term, then by program): reuse t and log(major)
                const uint32_t bc0 = SO(i), blen = SO(i + 1) - bc0;
                const double4 w4 = reinterpret_cast<const double4 *>(sW)[i];
                bool same = (blen == mB_len) && !noreuse;
                if (same) {
                    const uint2 a0 = sprog[bc0], a1 = blen > 1u ? sprog[bc0 + 1] : mB_w1, a2 = blen > 2u ? sprog[bc0 + 2] : mB_w2;
                    same = a0.x == mB_w0.x && a0.y == mB_w0.y && a1.x == mB_w1.x && a1.y == mB_w1.y && a2.x == mB_w2.x && a2.y == mB_w2.y;
                }
                if (same) {
                    acc += fma(w4.x, (w4.x != 0.0) ? mB_lm : 0.0, -fma(w4.y, mB_t + SNAQ_K_LOG3, w4.w));
                    continue;
                }
                const uint16_t *pc = reinterpret_cast<const uint16_t *>(sprog + bc0);
                const unsigned cw = pc[0];
                const int npairs = (int)(cw & 255u), nterms = (int)(cw >> 8);
                double R = 0.0;
                int k = 1;
#pragma unroll 1
                for (int j = 0; j < npairs; j++, k += 2) {
                    const unsigned wv = (unsigned)pc[k] | ((unsigned)pc[k + 1] << 16);
                    R += XLO(wv) - XHI(wv);
                }
                double Bn = 0.0, Bf = 0.0;
#pragma unroll 1
                for (int it = 0; it < nterms; it++) {
                    const unsigned th = pc[k];
                    const double gq = XLO(th >> 4);
                    if ((th & 3u) == SNAQ_T2) {
                        const unsigned wv = (unsigned)pc[k + 1] | ((unsigned)pc[k + 2] << 16);
                        k += 3;
                        if (noreuse || th != m2_k0 || wv != m2_k1) {
                            const double e = snaq_clamp10(XLO(wv) - XHI(wv));
                            const double gh = (th & 4u) ? 1.0 - gq : gq;
                            m2_v = snaq_clamp10(-snaq_log(1.0 - gh * (1.0 - snaq_exp(-e, M)), M));
                            m2_k0 = th; m2_k1 = wv;
                        }
                        R += m2_v;
                    } else {
                        const unsigned wv = (unsigned)pc[k + 1] | ((unsigned)pc[k + 2] << 16);
                        const unsigned kk0 = (th & ~4u) | ((unsigned)pc[k + 3] << 16);
                        const int nchain = pc[k + 4];
                        k += 5;
                        double ch = 0.0;
#pragma unroll 1
                        for (int j = 0; j < nchain; j++, k += 2) {
                            const unsigned wc = (unsigned)pc[k] | ((unsigned)pc[k + 1] << 16);
                            ch += XLO(wc) - XHI(wc);
                        }
                        ch = snaq_clamp10(ch);
                        if (noreuse || kk0 != m4_k0 || wv != m4_k1) {
                            const double clo = XLO(wv), chi = XHI(wv), ck = XLO((unsigned)pc[k - 2 * nchain - 2]);
                            const double ya = snaq_exp(-snaq_clamp10(clo), M), yb = snaq_exp(-snaq_clamp10(ck - chi), M);
                            const double ye = snaq_exp(-snaq_clamp10(chi - clo), M);
                            const double ga = 1.0 - gq, gb = gq;
                            m4_l4 = -snaq_log(gb * gb * yb + ga * gb * (3.0 - ye) + ga * ga * ya, M);
                            m4_k0 = kk0; m4_k1 = wv;
                        }
                        double l4 = m4_l4;
                        if (l4 < SNAQ_MINLEN) st |= SNAQ_S_NEGLEN;
                        l4 = snaq_clamp10(l4);
                        const double B = snaq_clamp10(ch + l4);
                        if (B < SNAQ_MINLEN) st |= SNAQ_S_NEGLEN;
                        if (th & 4u) Bf = B; else Bn = B;
                    }
                }
                double t = snaq_clamp10(Bn + R);
                if (t < SNAQ_MINLEN) st |= SNAQ_S_NEGLEN;
                t = snaq_clamp10(t + Bf);
                if (t < SNAQ_MINLEN) st |= SNAQ_S_NEGLEN;
                const double y = snaq_exp(-t, M);
                const double major = fma(SNAQ_K_M23, y, 1.0);
                if (major < 0.0) st |= SNAQ_S_NEGLEN;
                mB_lm = snaq_log(major, M);
                mB_t = t;
                if (blen <= 3u) {                                            // longer programs are not remembered
                    mB_len = blen;
                    mB_w0 = sprog[bc0];
                    mB_w1 = blen > 1u ? sprog[bc0 + 1] : make_uint2(0u, 0u);
                    mB_w2 = blen > 2u ? sprog[bc0 + 2] : make_uint2(0u, 0u);
                } else mB_len = 0xffffffffu;
                acc += fma(w4.x, (w4.x != 0.0) ? mB_lm : 0.0, -fma(w4.y, t + SNAQ_K_LOG3, w4.w));      // obs == 0 contributes 0 (src/pseudolik.jl:1398)
            }
        } else {
            // class C: 4-cycle quartets.  T5: [gslot, C_a, C_m, C_b] in one chunk; all three CFs are positive.  Sorted by these
            // four slots: the three logs are reused while they repeat.
            const XLane xf{xl};
            unsigned m5_k0 = 0xffffffffu, m5_k1 = 0xffffffffu;
            double m5_l0 = 0.0, m5_l1 = 0.0, m5_l2 = 0.0;
            const uint8_t *hd = qhdr + q0;
#pragma unroll 1
            for (int i = wbeg; i < wend; i++) {
                const uint2 w = sprog[SO(i)];
                const double4 w4 = reinterpret_cast<const double4 *>(sW)[i];
                if (SNAQ_HDR_KIND(hd[i]) == SNAQ_KIND_T5) {
                    if (noreuse || w.x != m5_k0 || w.y != m5_k1) {
                        const double g2 = XLO(w.x), ca = XHI(w.x), cm = XLO(w.y), cb = XHI(w.y);
                        const double g1 = 1.0 - g2;
                        const double y5 = snaq_exp(-snaq_clamp10(cm - ca), M), y6 = snaq_exp(-snaq_clamp10(cb - cm), M);
                        const double g13 = snaq_div3(g1), g23 = snaq_div3(g2);          // = g/3.0 bit for bit, no division
                        const double cf0 = g1 * (1.0 - 2.0 / 3.0 * y5) + g23 * y6;
                        const double cf1 = g13 * y5 + g2 * (1.0 - 2.0 / 3.0 * y6);
                        const double cf2 = g13 * y5 + g23 * y6;
                        m5_l0 = snaq_log_pos(cf0, M); m5_l1 = snaq_log_pos(cf1, M); m5_l2 = snaq_log_pos(cf2, M);
                        m5_k0 = w.x; m5_k1 = w.y;
                    }
                    double v = -w4.w;
                    v = fma(w4.x, m5_l0, v);
                    v = fma(w4.y, m5_l1, v);
                    v = fma(w4.z, m5_l2, v);
                    acc += v;
                } else {
                    const uint16_t *pc = reinterpret_cast<const uint16_t *>(sprog + SO(i));
                    double cf[3];
                    snaq_t5_cf(SNAQ_KIND_T5BD1, pc, xf, M, cf);
                    acc += snaq_loglik_t5(cf, sW + 4 * i, M, st, cr.pen ? cr.pen + 4 * (q0 + i) : nullptr);
                }
            }
        }

        const bool last_of_super = (cnext != cbi + 1);
        if (last_of_super) { red[warp * 32 + lane] = acc; acc = 0.0; }
        __syncthreads();                                                  // every warp is done with this stage and with soff
        if (last_of_super && warp == 0) {
            double s = 0.0;
#pragma unroll
            for (int w = 0; w < WARPS; w++) s += red[w * 32 + lane];
            partial[(size_t)(cbi / LANESP_SUPER) * Ppad + 32 * g + lane] = s;
        }
        cbi = cnext;
    }
#undef XLO
#undef XHI
    if (st) atomicOr(status, st);
}

/*
 * Run-length batched kernel (the default for P >= 16 on the plain table).
 *
 * The class-sorted table keeps quartets with the SAME program next to each other, and a quartet's expected CFs depend on
 * its program only.  set_network therefore also builds a run table (one program per run of identical programs: off_u /
 * prog_u) and from it a PACKED TILE STREAM: one self-contained record per tile of 64 tree-like (32 four-cycle) quartets,
 *     [ 1024 B per-quartet weights | 32 B descriptor (bit mask of the run heads, ...) | program offsets | programs ],
 * so that a tile arrives with ONE TMA bulk copy.  The kernel evaluates a program once per run and candidate and then
 * streams the run's per-quartet weights: two FMAs per (quartet, candidate) for tree-like quartets (w0 * log(major) and
 * w1 * (t1 + log 3)), three for 4-cycle quartets -- no compares, no offsets, no program loads per quartet.  The constants
 * sum_q W3 are added once by the reduction pass (negw3[1]).  Bad-diamond-I 4-cycles keep the per-quartet -1e15 penalty
 * semantics of src/pseudolik.jl:1394-1396 (snaq_loglik_t5).
 *
 * grid (gx, groups of 32 * CPL candidates); a lane holds CPL candidates (the weights are loaded once for all of them).
 * Every WARP owns a ring of RN_NST stages and walks whole SEGMENTS (tps consecutive tiles; tps depends on the table
 * only): one partial row per segment, summed in tile order by one warp, so a value depends neither on the grid nor on
 * the batch the candidate was evaluated in.
 */
#define RN_NST 3
#define RN_MIN_AVG_RUN 32   /* tables whose runs are shorter on average keep the per-quartet batched kernels */
#define RN_W_BYTES 1024
#define RN_DESC_AT RN_W_BYTES
#define RN_OFF_AT (RN_W_BYTES + 32)
#define RN_PROG_CHUNKS 64   /* program chunks a tile record can hold (else: read from the run table in global memory) */
#define RN_STAGE_BYTES (RN_OFF_AT + 272 + RN_PROG_CHUNKS * 8)        /* 2352: the largest record */
struct __align__(16) RunTile {
    unsigned long long mask;   // bit j: quartet q0 + j starts a run
    uint16_t n;                // number of quartets
    uint16_t flags;            // bit 0: 4-cycle tile (weights are W[q], 32 B); bit 1: programs are in the record
    uint16_t prog_at;          // byte offset of the programs inside the record
    uint16_t nruns;
    uint32_t q0;               // sorted position of the first quartet
    uint32_t pbeg;             // first program chunk of the tile's runs in prog_u (used when the programs are not in the record)
    uint32_t u0, pad;          // run of quartet q0
};
static_assert(sizeof(RunTile) == 32, "RunTile");
struct RunsMeta {
    uint32_t ntiles, nseg, tps, nul;
    int64_t nA;                // sorted positions [0,nA) additive, then type-2/4 terms, then (own tiles) 4-cycles
    int noreuse;
};

// pass 1 over the tiles: descriptor and record size (16-byte units); pass 2 (after the scan): offsets and programs
__global__ void __launch_bounds__(256) k_run_tiles(const uint32_t *__restrict__ flags, const uint32_t *__restrict__ uidx,
                                                   const uint32_t *__restrict__ off_u, int64_t nq, int64_t nAB, uint32_t ntAB,
                                                   uint32_t ntiles, RunTile *__restrict__ tdesc, uint32_t *__restrict__ size16) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntiles) return;
    RunTile d;
    const bool isC = t >= ntAB;
    const int64_t q0 = isC ? nAB + (int64_t)(t - ntAB) * RN_TQC : (int64_t)t * RN_TQ;
    const int64_t n = isC ? min((int64_t)RN_TQC, nq - q0) : min((int64_t)RN_TQ, nAB - q0);
    unsigned long long m = 0ull;
    for (int j = 0; j < (int)n; j++) m |= (unsigned long long)(flags[q0 + j] & 1u) << j;
    d.mask = m;
    d.u0 = uidx[q0] - 1u;
    d.nruns = (uint16_t)(__popcll(m >> 1) + 1);
    d.pbeg = off_u[d.u0];
    const uint32_t plen = off_u[d.u0 + d.nruns] - d.pbeg;
    const bool staged = plen <= (uint32_t)RN_PROG_CHUNKS;
    const uint32_t ob = (((uint32_t)d.nruns + 1u) * 4u + 15u) & ~15u;
    d.q0 = (uint32_t)q0; d.n = (uint16_t)n;
    d.prog_at = (uint16_t)(RN_OFF_AT + ob);
    d.flags = (uint16_t)((isC ? 1u : 0u) | (staged ? 2u : 0u));
    d.pad = 0u;
    tdesc[t] = d;
    size16[t] = (RN_OFF_AT + ob + (staged ? ((plen * 8u + 15u) & ~15u) : 0u)) / 16u;
}
__global__ void __launch_bounds__(256) k_tile_nruns(const RunTile *__restrict__ tdesc, uint32_t ntiles, uint32_t *__restrict__ out) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t <= ntiles) out[t] = t < ntiles ? (uint32_t)tdesc[t].nruns : 0u;
}
__global__ void __launch_bounds__(256) k_run_fill(const RunTile *__restrict__ tdesc, const uint32_t *__restrict__ rec_off,
                                                  const uint32_t *__restrict__ off_u, const uint2 *__restrict__ prog_u, uint32_t ntiles,
                                                  unsigned char *__restrict__ pack) {
    // one warp per tile
    const uint32_t t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    if (t >= ntiles) return;
    const RunTile d = tdesc[t];
    unsigned char *rec = pack + (size_t)rec_off[t] * 16;
    if (lane == 0) *reinterpret_cast<RunTile *>(rec + RN_DESC_AT) = d;
    uint32_t *roff = reinterpret_cast<uint32_t *>(rec + RN_OFF_AT);
    const uint32_t nslot = (d.prog_at - RN_OFF_AT) / 4u;
    for (uint32_t k = lane; k < nslot; k += 32u) roff[k] = k <= d.nruns ? off_u[d.u0 + k] - d.pbeg : 0u;   // relative to the tile's first program
    if (d.flags & 2u) {
        const uint32_t plen = off_u[d.u0 + d.nruns] - d.pbeg;
        uint2 *dst = reinterpret_cast<uint2 *>(rec + d.prog_at);
        for (uint32_t c = lane; c < ((plen + 1u) & ~1u); c += 32u) dst[c] = c < plen ? prog_u[d.pbeg + c] : make_uint2(0u, 0u);
    }
}

// ---- run-summed weights: W_u[u] = sum of the weights of run u's quartets (and the penalty data of snaq_loglik_t5).  Runs are very
// uneven (one quartet to hundreds of thousands), so the sum is formed over the TILES of the run table: k_wu_partials, one warp
// per tile, adds the tile's quartets per run (fixed shuffle tree); k_wu_combine, one warp per run, adds the run's consecutive
// partials with compensated sums and a fixed tree of error-free additions.  Deterministic, balanced, accurate to a rounding.
__global__ void __launch_bounds__(128) k_wu_partials(const RunTile *__restrict__ tdesc, const uint32_t *__restrict__ pidx,
                                                     const double4 *__restrict__ W, uint32_t ntiles, double *__restrict__ part) {
    const uint32_t t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    if (t >= ntiles) return;
    const RunTile d = tdesc[t];
    const unsigned long long m = d.mask | 1ull;
    double v[2][8];
    int r[2];
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const unsigned j = lane + 32u * k;
        r[k] = -1;
#pragma unroll
        for (int c = 0; c < 8; c++) v[k][c] = 0.0;
        if (j < d.n) {
            const double4 w = W[(size_t)d.q0 + j];
            r[k] = __popcll(m & ((2ull << j) - 1ull)) - 1;
            v[k][0] = w.x; v[k][1] = w.y; v[k][2] = w.z; v[k][3] = w.w;
            v[k][4] = w.x != 0.0 ? w.x * log(w.x / 100.0) : 0.0;
            v[k][5] = w.y != 0.0 ? w.y * log(w.y / 100.0) : 0.0;
            v[k][6] = w.z != 0.0 ? w.z * log(w.z / 100.0) : 0.0;
            v[k][7] = (w.x != 0.0 || w.y != 0.0 || w.z != 0.0) ? 1.0 : 0.0;
        }
    }
    double *out = part + (size_t)pidx[t] * 8;
    for (int rr = 0; rr < (int)d.nruns; rr++) {
#pragma unroll
        for (int c = 0; c < 8; c++) {
            double x = (r[0] == rr ? v[0][c] : 0.0) + (r[1] == rr ? v[1][c] : 0.0);
            for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
            if (lane == 0) out[(size_t)rr * 8 + c] = x;
        }
    }
}
__global__ void __launch_bounds__(128) k_wu_combine(const uint32_t *__restrict__ rstart, const RunTile *__restrict__ tdesc,
                                                    const uint32_t *__restrict__ pidx, const double *__restrict__ part, int64_t nu,
                                                    int64_t nq, int64_t nAB, uint32_t ntAB, double4 *__restrict__ Wu,
                                                    double4 *__restrict__ pen) {
    const int64_t u = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const unsigned lane = threadIdx.x & 31u;
    if (u >= nu) return;
    const int64_t i0 = rstart[u], i1 = (u + 1 < nu) ? (int64_t)rstart[u + 1] : nq;
    auto tile_of = [&](int64_t i) { return i < nAB ? (uint32_t)(i / RN_TQ) : ntAB + (uint32_t)((i - nAB) / RN_TQC); };
    const uint32_t t0 = tile_of(i0), t1 = tile_of(i1 - 1);
    const size_t first = (size_t)pidx[t0] + ((uint32_t)u - tdesc[t0].u0);
    const uint32_t cnt = t1 - t0 + 1;
    double s[8], c[8];
#pragma unroll
    for (int k = 0; k < 8; k++) { s[k] = 0.0; c[k] = 0.0; }
    for (uint32_t k = lane; k < cnt; k += 32u) {
        const double *p = part + (first + k) * 8;
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const double y = __dsub_rn(p[q], c[q]);
            const double tt = __dadd_rn(s[q], y);
            c[q] = __dsub_rn(__dsub_rn(tt, s[q]), y);
            s[q] = tt;
        }
    }
    double res[8];
#pragma unroll
    for (int q = 0; q < 8; q++) {
        double hi = s[q], lo = -c[q];
        for (int o = 16; o > 0; o >>= 1) {
            const double ohi = __shfl_xor_sync(0xffffffffu, hi, o), olo = __shfl_xor_sync(0xffffffffu, lo, o);
            // (a symmetric combination: both partners form the same pair)
            const double a = (lane & o) ? ohi : hi, b = (lane & o) ? hi : ohi;
            double ss, ee;
            two_sum(a, b, ss, ee);
            hi = ss;
            lo = __dadd_rn(__dadd_rn(lo, olo), ee);
        }
        res[q] = __dadd_rn(hi, lo);
    }
    if (lane == 0) {
        Wu[u] = make_double4(res[0], res[1], res[2], res[3]);
        pen[u] = make_double4(res[4], res[5], res[6], res[7]);
    }
}

// the candidates of a lane in the xf tile [row][CPL * 32]: element (row, c, lane)
struct XLaneS {
    const double *p; unsigned stride;
    __device__ __forceinline__ double operator()(unsigned s) const {
#ifdef SNAQ_BOUNDS
        SNAQ_BCHK((int)s < g_dbg_nfull);
#endif
        return p[s * stride];
    }
};

template <int CPL> struct RunConsts {       // what a run of identical programs contributes per candidate
    double cL[CPL], cT[CPL], c2[CPL];         // tree-like: log(major), t1 + log 3; 4-cycle: the three logs
    double cz[CPL][3];                        // bad diamond I: the three expected CFs
    bool bd1;
};
// Evaluation of the r-th run's program of a tile for the CPL candidates of this lane.  Out of line: it runs once per run,
// and keeping its temporaries out of the streaming loop's register allocation is what lets the loop hold its accumulators
template <int CPL>
__device__ __noinline__ void runs_compute(const RunTile &d, const uint32_t *soff, const uint2 *sprog, const uint2 *__restrict__ prog_u,
                                          const double *xl, const SnaqMath &M, const unsigned nul, const int64_t nA, const uint32_t r,
                                          const int64_t pos, RunConsts<CPL> &K, unsigned &st) {
    constexpr int TW = 32 * CPL;
    const bool staged = (d.flags & 2u) != 0, isC = (d.flags & 1u) != 0;
    auto chunk = [&](uint32_t c) { return staged ? sprog[c] : __ldg(prog_u + d.pbeg + c); };

                SNAQ_BCHK(r < d.nruns);
                const uint32_t c0 = soff[r], c1 = soff[r + 1];
                SNAQ_BCHK(c0 < c1 && (!staged || d.prog_at + c1 * 8u <= (unsigned)RN_STAGE_BYTES));
                if (isC) {
                    const uint2 w = chunk(c0);
                    K.bd1 = (w.y & 0xffffu) == nul;
    #pragma unroll
                    for (int c = 0; c < CPL; c++) {
                        const XLaneS xf{xl + 32 * c, (unsigned)TW};
                        if (!K.bd1) {
                            const double g2 = xf(w.x & 0xffffu), ca = xf(w.x >> 16), cm = xf(w.y & 0xffffu), cb = xf(w.y >> 16);
                            const double g1 = 1.0 - g2;
                            const double y5 = snaq_exp(-snaq_clamp10(cm - ca), M), y6 = snaq_exp(-snaq_clamp10(cb - cm), M);
                            const double g13 = snaq_div3(g1), g23 = snaq_div3(g2);          // = g/3.0 bit for bit, no division
                            const double cf0 = g1 * (1.0 - 2.0 / 3.0 * y5) + g23 * y6;
                            const double cf1 = g13 * y5 + g2 * (1.0 - 2.0 / 3.0 * y6);
                            const double cf2 = g13 * y5 + g23 * y6;
                            K.cL[c] = snaq_log_pos(cf0, M); K.cT[c] = snaq_log_pos(cf1, M); K.c2[c] = snaq_log_pos(cf2, M);
                        } else {
                            const uint16_t pcw[4] = {(uint16_t)(w.x & 0xffffu), (uint16_t)(w.x >> 16), (uint16_t)nul, (uint16_t)nul};
                            snaq_t5_cf(SNAQ_KIND_T5BD1, pcw, xf, M, K.cz[c]);
                        }
                    }
                } else if (pos >= nA) {
                    const uint16_t *pc = staged ? reinterpret_cast<const uint16_t *>(sprog + c0)
                                                : reinterpret_cast<const uint16_t *>(prog_u + d.pbeg + c0);
    #pragma unroll
                    for (int c = 0; c < CPL; c++) {
                        const XLaneS xf{xl + 32 * c, (unsigned)TW};
                        const double tt = snaq_tree_terms_t1(pc, xf, M, st);
                        const double major = fma(SNAQ_K_M23, snaq_exp(-tt, M), 1.0);
                        if (major < 0.0) st |= SNAQ_S_NEGLEN;
                        K.cL[c] = snaq_log(major, M);
                        K.cT[c] = tt + SNAQ_K_LOG3;
                    }
                } else {
                    double t0[CPL], t1[CPL];
    #pragma unroll
                    for (int c = 0; c < CPL; c++) { t0[c] = 0.0; t1[c] = 0.0; }
    #pragma unroll 1
                    for (uint32_t cc = c0; cc < c1; ++cc) {
                        const uint2 w = chunk(cc);
    #pragma unroll
                        for (int c = 0; c < CPL; c++) {
                            const XLaneS xf{xl + 32 * c, (unsigned)TW};
                            t0[c] += xf(w.x & 0xffffu) - xf(w.x >> 16);
                            t1[c] += xf(w.y & 0xffffu) - xf(w.y >> 16);
                        }
                    }
    #pragma unroll
                    for (int c = 0; c < CPL; c++) {
                        const double tt = snaq_clamp10(t0[c] + t1[c]);
                        K.cL[c] = snaq_log_pos(fma(SNAQ_K_M23, snaq_exp(-tt, M), 1.0), M);
                        K.cT[c] = tt + SNAQ_K_LOG3;
                    }
                }
            }

template <int WARPS, int CPL>
__global__ void __launch_bounds__(WARPS * 32, (WARPS == 8 && CPL == 1) ? 2 : 1) k_eval_runs(const unsigned char *__restrict__ pack, const uint32_t *__restrict__ rec_off,
                                                        const uint2 *__restrict__ prog_u, const RunsMeta rm,
                                                        const double *__restrict__ XFT, const int n_full,
                                                        double *__restrict__ partial, const int Ppad, uint32_t *__restrict__ status) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int TW = 32 * CPL;                                                     // candidates per block = tile width
    double *mt = reinterpret_cast<double *>(smem);                                   // 384 B
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + 384);                       // [0] xf tile, [1 + warp * NST + s]
    constexpr size_t BAR_BYTES = ((1 + WARPS * RN_NST) * 8 + 127) & ~size_t(127);
    double *xs = reinterpret_cast<double *>(smem + 384 + BAR_BYTES);
    unsigned char *ring = smem + 384 + BAR_BYTES + (size_t)n_full * TW * 8;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = blockIdx.y;
    const unsigned gw = blockIdx.x * WARPS + warp, nwg = gridDim.x * WARPS;          // this warp among the group's warps
    uint64_t *wbar = bars + 1 + warp * RN_NST;
    unsigned char *wring = ring + (size_t)warp * RN_NST * RN_STAGE_BYTES;
    const bool noreuse = rm.noreuse != 0;
    const unsigned nul = rm.nul;
    // this warp's segments: gw, gw + nwg, ... of tps consecutive tiles each; the request side runs RN_NST tiles ahead
    unsigned i_seg = gw, i_t = gw * rm.tps, i_end = min(i_t + rm.tps, rm.ntiles);
    auto i_advance = [&]() {
        if (++i_t >= i_end) { i_seg += nwg; i_t = i_seg * rm.tps; i_end = min(i_t + rm.tps, rm.ntiles); }
    };
    auto issue = [&](unsigned o0, unsigned o1, unsigned sg) {                        // lane 0 only: one bulk copy per tile
        const unsigned bytes = (o1 - o0) * 16u;
        mbar_expect_tx(&wbar[sg], bytes);
        tma_load_1d(wring + sg * RN_STAGE_BYTES, pack + (size_t)o0 * 16, bytes, &wbar[sg]);
    };
    if (tid == 0) mbar_init(&bars[0], 1);
    if (lane == 0) {
        for (int k = 0; k < RN_NST; k++) mbar_init(&wbar[k], 1);
    }
    __syncthreads();
    grid_launch_dependents();
    if (tid == 0) {
        mbar_expect_tx(&bars[0], (unsigned)n_full * TW * 8u);
        grid_dependency_wait();                                  // the tile is k_expand_x's output (same step)
        tma_load_1d(xs, XFT + (size_t)g * n_full * TW, (unsigned)n_full * TW * 8u, &bars[0]);
    }
    // the warp's first tiles are requested before anything else is waited for
    for (unsigned k = 0; k < RN_NST && i_seg < rm.nseg; k++) {
        if (lane == 0) issue(__ldg(rec_off + i_t), __ldg(rec_off + i_t + 1), k);
        i_advance();
    }
    unsigned no0 = 0, no1 = 0;                                                       // the record of the next tile to request
    if (i_seg < rm.nseg) { no0 = __ldg(rec_off + i_t); no1 = __ldg(rec_off + i_t + 1); }
    __syncwarp();
    const SnaqMath M = load_math_tables(mt, tid);
    unsigned st = 0;
    mbar_wait(&bars[0], 0);
    __syncthreads();
    const double *xl = xs + lane;
    unsigned sg = 0, phase = 0;
    // segment sums: w0*L (a0, b0, e0, f0), w1*T (a1, b1, e1, f1), 4-cycles.  Four independent accumulators per product
    // and candidate: the FP64 pipe has a long dependent-issue latency and a block has few warps (the xf tile is large)
    constexpr int NA = 4;                                       // accumulators per product and candidate
    double aL[CPL][NA], aT[CPL][NA], cacc[CPL];
    double cL[CPL], cT[CPL], c2[CPL];                           // the current run's constants: tree-like (L, T); 4-cycle (l0, l1, l2)
    RunConsts<CPL> K;
    bool cBD1 = false;
#pragma unroll
    for (int c = 0; c < CPL; c++) {
#pragma unroll
        for (int q = 0; q < NA; q++) { aL[c][q] = 0.0; aT[c][q] = 0.0; }
        cacc[c] = 0.0;
        cL[c] = cT[c] = c2[c] = 0.0;
    }
#pragma unroll 1
    for (unsigned seg = gw; seg < rm.nseg; seg += nwg) {
        const unsigned t0 = seg * rm.tps, tend = min(t0 + rm.tps, rm.ntiles);
        bool seg_first = true;
#pragma unroll 1
        for (unsigned t = t0; t < tend; t++) {
            mbar_wait(&wbar[sg], phase);
            const unsigned char *stg = wring + sg * RN_STAGE_BYTES;
            const RunTile d = *reinterpret_cast<const RunTile *>(stg + RN_DESC_AT);
            SNAQ_BCHK(d.n >= 1 && d.n <= RN_TQ && d.nruns >= 1 && d.nruns <= d.n && d.prog_at >= RN_OFF_AT + 16 && d.prog_at <= RN_OFF_AT + 272);
            const uint32_t *soff = reinterpret_cast<const uint32_t *>(stg + RN_OFF_AT);
            const uint2 *sprog = reinterpret_cast<const uint2 *>(stg + d.prog_at);
            const bool isC = (d.flags & 1u) != 0;
            auto compute = [&](uint32_t r, int64_t pos) {
                runs_compute<CPL>(d, soff, sprog, prog_u, xl, M, nul, rm.nA, r, pos, K, st);
#pragma unroll
                for (int c = 0; c < CPL; c++) { cL[c] = K.cL[c]; cT[c] = K.cT[c]; c2[c] = K.c2[c]; }
                cBD1 = K.bd1;
            };
            const double2 *sw = reinterpret_cast<const double2 *>(stg);
            if (!isC && !noreuse && (d.mask >> 1) == 0ull && d.n == RN_TQ) {
                // the whole tile is one run (the common case on large tables): straight-line, no loop control
                if ((d.mask & 1ull) || seg_first) compute(0u, (int64_t)d.q0);
    #pragma unroll
                for (int i = 0; i < RN_TQ; i += NA) {
                    double2 w[NA];
    #pragma unroll
                    for (int q = 0; q < NA; q++) w[q] = sw[i + q];
    #pragma unroll
                    for (int q = 0; q < NA; q++) {
    #pragma unroll
                        for (int c = 0; c < CPL; c++) { aL[c][q] = fma(w[q].x, cL[c], aL[c][q]); aT[c][q] = fma(w[q].y, cT[c], aT[c][q]); }
                    }
                }
            } else {
                unsigned j = 0, r = 0;
                const unsigned n = d.n;
    #pragma unroll 1
                while (j < n) {
                    const unsigned long long rest = (d.mask >> j) >> 1;
                    const unsigned jend = rest ? min(j + (unsigned)__ffsll((long long)rest), n) : n;
                    if (j > 0) r++;
                    if (j > 0 || (d.mask & 1ull) || seg_first) compute(r, (int64_t)d.q0 + j);
                    if (!isC) {
                        if (!noreuse) {
                            unsigned i = j;
    #pragma unroll 1
                            for (; i + 3 < jend; i += 4) {
                                const double2 wa = sw[i], wb = sw[i + 1], wc = sw[i + 2], wd = sw[i + 3];
    #pragma unroll
                                for (int c = 0; c < CPL; c++) {
                                    aL[c][0] = fma(wa.x, cL[c], aL[c][0]); aT[c][0] = fma(wa.y, cT[c], aT[c][0]);
                                    aL[c][1] = fma(wb.x, cL[c], aL[c][1]); aT[c][1] = fma(wb.y, cT[c], aT[c][1]);
                                    aL[c][2 % NA] = fma(wc.x, cL[c], aL[c][2 % NA]); aT[c][2 % NA] = fma(wc.y, cT[c], aT[c][2 % NA]);
                                    aL[c][3 % NA] = fma(wd.x, cL[c], aL[c][3 % NA]); aT[c][3 % NA] = fma(wd.y, cT[c], aT[c][3 % NA]);
                                }
                            }
    #pragma unroll 1
                            for (; i < jend; i++) {
                                const double2 wa = sw[i];
    #pragma unroll
                                for (int c = 0; c < CPL; c++) { aL[c][0] = fma(wa.x, cL[c], aL[c][0]); aT[c][0] = fma(wa.y, cT[c], aT[c][0]); }
                            }
                        } else {
                            // measurement aid (SNAQ_REUSE=0): the program is evaluated again for every quartet
    #pragma unroll 1
                            for (unsigned i = j; i < jend; i++) {
                                compute(r, (int64_t)d.q0 + j);
                                const double2 wa = sw[i];
    #pragma unroll
                                for (int c = 0; c < CPL; c++) { aL[c][0] = fma(wa.x, cL[c], aL[c][0]); aT[c][0] = fma(wa.y, cT[c], aT[c][0]); }
                            }
                        }
                    } else {
                        const double4 *sw4 = reinterpret_cast<const double4 *>(stg);
    #pragma unroll 1
                        for (unsigned i = j; i < jend; i++) {
                            if (noreuse) compute(r, (int64_t)d.q0 + j);
                            const double4 w4 = sw4[i];
                            if (!cBD1) {
    #pragma unroll
                                for (int c = 0; c < CPL; c++) {
                                    aL[c][0] = fma(w4.x, cL[c], aL[c][0]);
                                    aL[c][1] = fma(w4.y, cT[c], aL[c][1]);
                                    cacc[c] = fma(w4.z, c2[c], cacc[c]);
                                }
                            } else {
                                const double wq[4] = {w4.x, w4.y, w4.z, w4.w};
    #pragma unroll
                                for (int c = 0; c < CPL; c++) cacc[c] += snaq_loglik_t5(K.cz[c], wq, M, st, nullptr);   // its own W3 and the per-quartet penalty
                            }
                        }
                    }
                    j = jend;
                }
            }
            seg_first = false;
            __syncwarp();                                              // every lane is done with this stage
            if (i_seg < rm.nseg) {
                if (lane == 0) issue(no0, no1, sg);
                i_advance();
                if (i_seg < rm.nseg) { no0 = __ldg(rec_off + i_t); no1 = __ldg(rec_off + i_t + 1); }   // consumed one tile later
            }
            if (++sg == RN_NST) { sg = 0; phase ^= 1u; }
        }
        // the segment is complete: one partial row per segment
#pragma unroll
        for (int c = 0; c < CPL; c++) {
            double sl = 0.0, stt = 0.0;
#pragma unroll
            for (int q = 0; q < NA; q++) { sl += aL[c][q]; stt += aT[c][q]; aL[c][q] = 0.0; aT[c][q] = 0.0; }
            partial[(size_t)seg * Ppad + TW * g + 32 * c + lane] = (sl - stt) + cacc[c];
            cacc[c] = 0.0;
        }
    }
    if (st) atomicOr(status, st);
}

/*
 * Thread per quartet (class-sorted order), at most PC candidates per launch: the HBM-bound path of a single
 * objective call (P = 1) and of small batches.  One persistent launch, no block-level synchronisation in the
 * main loops: every WARP owns a ring of QD_NST shared-memory stages fed by TMA bulk copies (cp.async.bulk +
 * one mbarrier per stage), so loads stay in flight without holding registers and a warp only ever waits for
 * its own data.
 *   entry    : lane 0 of every warp starts the bulk copies of the warp's first QD_NST tiles, THEN
 *   prologue : the block expands x into xf (snaq_expand_row) in shared memory, PC rows
 *   region G : everything that is not a short additive program (longer additive programs, type-2/4 terms,
 *              4-cycles) through the generic evaluator, 32 consecutive quartets per warp step; the warp
 *              stages the contiguous program span of its quartets in shared memory with coalesced loads
 *   region A : tiles of QD_TQ = 64 additive quartets, strided over all warps of the grid.  Tile of
 *              A1 [0,nA1): 16 B weight pair + one 8 B program chunk per quartet (program i = chunk i);
 *              A2 [nA1,nA1+nA2): 16 B + two chunks (program j at chunk a2base + 2 j; a2base is even so
 *              that the bulk copies are 16-byte aligned).  Each tile = two bulk copies (1 KB + 0.5-1 KB).
 *   epilogue : thread-private sums -> fixed-order block reduction -> partial[block][p]; the last block to
 *              finish (atomic ticket) adds all partials in block order: out[p] = -(sum) + sum of the A1/A2
 *              constants W3 (negw3 = -sum W3, precomputed with the weights).  Bit-reproducible.
 * OUT: per-quartet results (P = 1) at the quartet's ORIGINAL index (expcf [nq][3] in data order, logpl [nq],
 * deltacf [nq]); every quartet goes through the generic evaluator.
 */
#define QD_SPAN 96     /* program chunks a warp can stage for its 32 quartets (else: read from global memory) */
#define QD_TQ 64       /* additive quartets per warp tile */
#define QD_NST 3       /* stages per warp */
#define QD_STAGE_BYTES (QD_TQ * 32)
template <int PC, bool OUT>
__global__ void __launch_bounds__(256, OUT ? 2 : 3)
k_eval_direct(SnaqTables T, const uint32_t *__restrict__ off, const uint2 *__restrict__ prog2, const uint8_t *__restrict__ qhdr,
              const uint32_t *__restrict__ perm, const double2 *__restrict__ WA, const double4 *__restrict__ W,
              const double *__restrict__ obs, uint32_t nA1, uint32_t nA2, uint32_t a2base, uint32_t nQ,
              const double *__restrict__ X, const double *__restrict__ xconst, int nh, int nt, int P,
              double *__restrict__ partial, double *__restrict__ out, const double *__restrict__ negw3,
              unsigned int *__restrict__ ticket, double *__restrict__ expcf, double *__restrict__ logpl,
              double *__restrict__ deltacf, uint32_t *__restrict__ status, unsigned long long *__restrict__ trace,
              uint32_t *__restrict__ status_out, const int use_xarg, const double *__restrict__ pen, const __grid_constant__ PeerBox box,
              const __grid_constant__ XArg xarg) {
    constexpr int BLOCK = 256;
    extern __shared__ __align__(128) unsigned char smem[];
    double *mt = reinterpret_cast<double *>(smem);                        // 48 doubles, 128-byte aligned
    // optional phase timeline (diagnostics, SNAQ_TRACE=1): globaltimer of thread 0 at entry / after the prologue /
    // after region G / after region A / at exit
    auto stamp = [&](int k) {
        if (trace && threadIdx.x == 0) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            trace[(size_t)blockIdx.x * 8 + k] = t;
        }
    };
    stamp(0);
    static_assert(8 * QD_NST * 8 <= 256, "mbarrier area");
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + 384);            // [8 warps][QD_NST] mbarriers (<= 256 B)
    double *red = reinterpret_cast<double *>(smem + 640);                 // [8][PC] warp sums
    uint2 *span = reinterpret_cast<uint2 *>(smem + 640 + 8 * PC * 8);     // [8 warps][QD_SPAN] staged programs
    unsigned char *ring = smem + 640 + 8 * PC * 8 + 8 * QD_SPAN * 8;      // [8 warps][QD_NST][QD_STAGE_BYTES]
    double *xs = reinterpret_cast<double *>(ring + (OUT ? 0 : 8 * QD_NST * QD_STAGE_BYTES));   // [PC][n_full]
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const int n_full = T.n_full;
    double *xraw = xs + PC * n_full;                                      // [PC][nparams] the candidates themselves
    unsigned st = 0;
    const unsigned gwarp = blockIdx.x * (BLOCK / 32) + warp, nwarps = gridDim.x * (BLOCK / 32);
    const unsigned tA1 = (nA1 + QD_TQ - 1) / QD_TQ, tA2 = (nA2 + QD_TQ - 1) / QD_TQ, tA = tA1 + tA2;
    uint64_t *wbar = bars + warp * QD_NST;
    unsigned char *wring = ring + (size_t)warp * QD_NST * QD_STAGE_BYTES;
    // a tile's two bulk copies: weight pairs to the first 2 KB of the stage, program chunks behind them
    auto issue = [&](unsigned t, unsigned s) {                            // lane 0 only
        unsigned char *dst = wring + s * QD_STAGE_BYTES;
        if (t < tA1) {
            const unsigned i0 = t * QD_TQ, n = min((unsigned)QD_TQ, nA1 - i0);
            const unsigned pb = ((n + 1u) & ~1u) * 8u;
            mbar_expect_tx(&wbar[s], n * 16u + pb);
            tma_load_1d(dst, WA + i0, n * 16u, &wbar[s]);
            tma_load_1d(dst + QD_TQ * 16, prog2 + i0, pb, &wbar[s]);
        } else {
            const unsigned j0 = (t - tA1) * QD_TQ, n = min((unsigned)QD_TQ, nA2 - j0);
            mbar_expect_tx(&wbar[s], n * 32u);
            tma_load_1d(dst, WA + nA1 + j0, n * 16u, &wbar[s]);
            tma_load_1d(dst + QD_TQ * 16, prog2 + a2base + 2u * j0, n * 16u, &wbar[s]);
        }
    };
    if (!OUT) {
        if (lane == 0) {
            for (int k = 0; k < QD_NST; k++) mbar_init(&wbar[k], 1);
            for (unsigned k = 0; k < QD_NST; k++)
                if (gwarp + k * nwarps < tA) issue(gwarp + k * nwarps, k);
        }
        __syncwarp();
    }
    const SnaqMath M = load_math_tables(mt, (int)tid);
    // prologue: the candidates (kernel argument or device memory) -> shared memory, with the bounds check of update!;
    // then the expanded parameter vectors; unused rows repeat the last candidate
    for (int idx = (int)tid; idx < PC * T.nparams; idx += BLOCK) {
        const int p = min(idx / T.nparams, P - 1), row = idx % T.nparams;
        const double v = use_xarg ? xarg.v[p * T.nparams + row] : X[(size_t)p * T.nparams + row];
        st |= range_check(v, (row >= nh && row < nh + nt) ? 1.0e300 : 1.0);
        xraw[idx] = v;
    }
    __syncthreads();
    for (int idx = (int)tid; idx < PC * n_full; idx += BLOCK) {
        const int p = idx / n_full, row = idx % n_full;
        xs[idx] = snaq_expand_row(T, M, xraw + p * T.nparams, xconst, row);
    }
    __syncthreads();
    stamp(1);
    double acc[PC];
#pragma unroll
    for (int p = 0; p < PC; p++) acc[p] = 0.0;
    // ---- region G (everything when OUT) ---------------------------------------------------------------
    {
        const unsigned g0 = OUT ? 0u : nA1 + nA2;
        uint2 *wspan = span + warp * QD_SPAN;
#pragma unroll 1
        for (unsigned ib = g0 + gwarp * 32u; ib < nQ; ib += nwarps * 32u) {
            const unsigned i = ib + lane;
            const bool active = i < nQ;
            const unsigned ilast = min(ib + 32u, nQ);
            const uint32_t o0 = active ? __ldg(off + i) : 0u, o1 = active ? __ldg(off + i + 1) : 0u;
            const unsigned hdr = active ? (unsigned)qhdr[i] : 0u;
            const double4 w4 = active ? W[i] : make_double4(0, 0, 0, 0);
            const uint32_t s0 = __shfl_sync(0xffffffffu, o0, 0);
            const uint32_t s1 = __ldg(off + ilast);
            const bool staged = (s1 - s0) <= (uint32_t)QD_SPAN;
            __syncwarp();
            if (staged) for (uint32_t c = s0 + lane; c < s1; c += 32u) wspan[c - s0] = __ldg(prog2 + c);
            __syncwarp();
            if (active) {
                const double w[4] = {w4.x, w4.y, w4.z, w4.w};
                const uint16_t *pc = reinterpret_cast<const uint16_t *>(staged ? wspan + (o0 - s0) : prog2 + o0);
                const int nwords = (int)(o1 - o0) * 4;
#pragma unroll 1
                for (int p = 0; p < PC; p++) {
                    if (p < P) {
                        const XRow xf{xs + p * n_full};
                        double cf[3], t1;
                        snaq_eval_quartet(hdr, pc, nwords, xf, M, cf, &t1, st);
                        const double v = snaq_quartet_loglik(SNAQ_HDR_KIND(hdr), cf, t1, w, M, st, (pen && !OUT) ? pen + 4 * (size_t)i : nullptr);
#pragma unroll
                        for (int k = 0; k < PC; k++) if (k == p) acc[k] += v;     // keeps acc[] in registers
                        if (OUT) {
                            const size_t q = perm[i];
                            double e[3];
                            snaq_data_order(hdr, cf, e);
                            if (expcf) { expcf[3 * q] = e[0]; expcf[3 * q + 1] = e[1]; expcf[3 * q + 2] = e[2]; }
                            if (logpl) logpl[q] = v;
                            if (deltacf) {   // calculateDeltaCF: src/pseudolik.jl:491-498 (0 for unsampled quartets, :514)
                                const bool smp = (w4.x != 0.0 || w4.y != 0.0 || w4.z != 0.0 || w4.w != 0.0);
                                const double d = fabs(obs[3 * q] - e[0]) + fabs(obs[3 * q + 1] - e[1]) + fabs(obs[3 * q + 2] - e[2]);
                                deltacf[q] = smp ? d : 0.0;
                            }
                        }
                    }
                }
            }
        }
    }
    stamp(2);
    // ---- region A: warp-private TMA ring ------------------------------------------------------------------
    if (!OUT) {
        const unsigned nullpair = (unsigned)T.n_xe * 0x10001u;
        auto tree_term = [&](double t, const double2 &w) {                // one additive quartet, one candidate
            const double t1 = snaq_clamp10(t);
            const double y = snaq_exp(-t1, M);
            const double major = fma(SNAQ_K_M23, y, 1.0);
            return fma(w.x, snaq_log_pos(major, M), -w.y * (t1 + SNAQ_K_LOG3));
        };
        unsigned s = 0, phase = 0;
#pragma unroll 1
        for (unsigned t = gwarp; t < tA; t += nwarps) {
            mbar_wait(&wbar[s], phase);
            const unsigned char *stg = wring + s * QD_STAGE_BYTES;
            const double2 *sw = reinterpret_cast<const double2 *>(stg);
            if (t < tA1) {
                const unsigned n = min((unsigned)QD_TQ, nA1 - t * QD_TQ);
                const uint2 *sp = reinterpret_cast<const uint2 *>(stg + QD_TQ * 16);
#pragma unroll
                for (int k = 0; k < QD_TQ / 32; k++) {
                    const unsigned idx = (unsigned)k * 32u + lane;
                    const bool ok = idx < n;
                    const double2 w = ok ? sw[idx] : make_double2(0.0, 0.0);
                    const uint2 c = ok ? sp[idx] : make_uint2(nullpair, nullpair);
#pragma unroll
                    for (int p = 0; p < PC; p++) {
                        const double *xp = xs + p * n_full;
                        const double tt = (xp[c.x & 0xffffu] - xp[c.x >> 16]) + (xp[c.y & 0xffffu] - xp[c.y >> 16]);
                        acc[p] += tree_term(tt, w);
                    }
                }
            } else {
                const unsigned n = min((unsigned)QD_TQ, nA2 - (t - tA1) * QD_TQ);
                const uint4 *sp = reinterpret_cast<const uint4 *>(stg + QD_TQ * 16);
#pragma unroll
                for (int k = 0; k < QD_TQ / 32; k++) {
                    const unsigned idx = (unsigned)k * 32u + lane;
                    const bool ok = idx < n;
                    const double2 w = ok ? sw[idx] : make_double2(0.0, 0.0);
                    const uint4 c = ok ? sp[idx] : make_uint4(nullpair, nullpair, nullpair, nullpair);
#pragma unroll
                    for (int p = 0; p < PC; p++) {
                        const double *xp = xs + p * n_full;
                        double ta = xp[c.x & 0xffffu] - xp[c.x >> 16], tb = xp[c.y & 0xffffu] - xp[c.y >> 16];
                        ta += xp[c.z & 0xffffu] - xp[c.z >> 16];
                        tb += xp[c.w & 0xffffu] - xp[c.w >> 16];
                        acc[p] += tree_term(ta + tb, w);
                    }
                }
            }
            __syncwarp();                                                  // every lane is done with stage s
            if (lane == 0 && t + QD_NST * nwarps < tA) issue(t + QD_NST * nwarps, s);
            if (++s == QD_NST) { s = 0; phase ^= 1u; }
        }
    }
    if (st) atomicOr(status, st);
    stamp(3);
    // ---- epilogue: fixed-order block reduction, then the last block adds the partials in block order -----
#pragma unroll
    for (int p = 0; p < PC; p++) {
        double v = acc[p];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (lane == 0) red[warp * PC + p] = v;
    }
    __syncthreads();
    if ((int)tid < PC) {
        double ssum = 0.0;
#pragma unroll
        for (int k = 0; k < BLOCK / 32; k++) ssum += red[k * PC + tid];
        partial[(size_t)blockIdx.x * PC + tid] = ssum;
    }
    __threadfence();
    __shared__ unsigned int s_last;
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1) ? 1u : 0u;
    __syncthreads();
    if (s_last) {
        __threadfence();
        const double cst = OUT ? 0.0 : __ldcg(negw3);
        __shared__ double s_val[4];
        const bool fused = !OUT && box.seq != 0ull;        // quartet-sharded: the all-reduce over the ranks happens here
        for (int p = 0; p < P; p++) {
            double sp = 0.0;
            for (unsigned b = tid; b < gridDim.x; b += BLOCK) sp += __ldcg(partial + (size_t)b * PC + p);
            for (int o = 16; o > 0; o >>= 1) sp += __shfl_down_sync(0xffffffffu, sp, o);
            __syncthreads();
            if (lane == 0) red[warp] = sp;
            __syncthreads();
            if (tid == 0) {
                double tot = 0.0;
#pragma unroll
                for (int k = 0; k < BLOCK / 32; k++) tot += red[k];
                if (fused) s_val[p & 3] = -tot - cst;
                else out[p] = -tot - cst;
            }
        }
        if (fused) {
            __syncthreads();
            const unsigned slot = (unsigned)(box.seq & 1ull);
            if ((int)tid < box.world) {
                // deliver this rank's partial sums to rank `tid` (plain stores through the peer mapping), then the flag
                volatile PeerMail *m = box.peer[tid] + (size_t)slot * box.world + box.rank;
                for (int p = 0; p < P; p++) m->v[p] = s_val[p];
                __threadfence_system();
                m->seq = box.seq;
                // and wait for rank `tid`'s entry in our own box
                const volatile PeerMail *in = box.peer[box.rank] + (size_t)slot * box.world + tid;
                unsigned long long t0;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
                while (in->seq != box.seq) {
                    unsigned long long t1;
                    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                    if (t1 - t0 > 4000000000ull) { atomicOr(status, SNAQ_S_COMM); break; }      // 4 s: a peer is gone
                }
                __threadfence_system();
            }
            __syncthreads();
            if ((int)tid < P) {
                const volatile PeerMail *in = box.peer[box.rank] + (size_t)slot * box.world;
                double tot = 0.0;
                for (int r = 0; r < box.world; r++) tot += in[r].v[tid];           // rank order: the same bits on every rank
                out[tid] = tot;
            }
            __threadfence_system();
            __syncthreads();
        }
        if (tid == 0) {
            *ticket = 0u;
            if (status_out) {                            // every block's bits are in by now; the host polls for the DONE bit
                const uint32_t st = *reinterpret_cast<volatile uint32_t *>(status);
                __threadfence_system();                  // out[] (same thread, page-locked host memory) lands before the flag
                *reinterpret_cast<volatile uint32_t *>(status_out) = st | SNAQ_S_DONE;
            }
        }
    }
    stamp(4);
}

/*
 * Objective AND gradient in one pass (snaq_eval_grad): thread per quartet, one parameter vector.
 * Every quartet adds d v / d xf[slot] for the slots its program reads into per-block bins in shared
 * memory.  The bins are 64-bit fixed point (2^-32): integer atomics make the sums independent of the order
 * in which threads arrive, so the result is bit-reproducible.  A 64-bit shared-memory atomic add is a
 * compare-and-swap loop on this architecture, so a bin is two 32-bit words updated by two native 32-bit
 * atomic adds, the carry of the low word (known from the value the first add returns) folded into the second.
 * A block's bins are exact in double, and the last block adds the per-block rows in block order.  The chain rule through the expansion x -> xf is
 * snaq_grad_backprop on the host (a few hundred rows).
 */
// One gradient bin = 96-bit two's-complement fixed point (2^-32) held as three 32-bit words in shared memory (a 64-bit
// shared-memory atomic is a compare-and-swap loop on sm_100; 32-bit adds are native).  Adds the sign-extended 64-bit q:
// the carry of the low word is known from the value its add returns and is folded into the add of the middle word,
// whose own carry (rare) goes to the top word.  Range +-2^63 value units per block and slot: it cannot wrap, whatever
// the table size and however close an expected CF comes to 0 (a single quartet's derivative is below 2^31 value units,
// its weights being at most 100 and every denominator at least exp(-10) / 3).
__device__ __forceinline__ void bin_add96(unsigned *lo, unsigned *hi, unsigned *ext, unsigned slot, long long q) {
    const unsigned ql = (unsigned)q, qh = (unsigned)((unsigned long long)q >> 32), qe = q < 0 ? 0xffffffffu : 0u;
    const unsigned old = atomicAdd(&lo[slot], ql);
    const unsigned inc = qh + ((old + ql < old) ? 1u : 0u);
    const unsigned c2a = (inc < qh) ? 1u : 0u;
    const unsigned oldh = atomicAdd(&hi[slot], inc);
    const unsigned e = qe + c2a + ((oldh + inc < oldh) ? 1u : 0u);
    if (e) atomicAdd(&ext[slot], e);
}
// the same for a contribution of any size (|v| < 2^63 value units): integer part and 2^-32 fraction are formed separately.
// Needed for the deduplicated table, where one entry stands for a whole run of quartets (its weights are sums) and a
// single contribution can exceed the 2^31 value units that fit the 64-bit form
__device__ __forceinline__ void bin_add96_wide(unsigned *lo, unsigned *hi, unsigned *ext, unsigned slot, double v) {
    const long long I = __double2ll_rn(v);
    const long long F = __double2ll_rn((v - (double)I) * 4294967296.0);      // |F| <= 2^31
    const unsigned fl = (unsigned)F;
    const unsigned old = atomicAdd(&lo[slot], fl);
    const long long U = I + (F >> 32) + ((old + fl < old) ? 1ll : 0ll);       // upper 64 bits of (I << 32) + F, plus the carry
    const unsigned uh = (unsigned)U, ue = (unsigned)((unsigned long long)U >> 32);
    const unsigned oldh = atomicAdd(&hi[slot], uh);
    const unsigned e = ue + ((oldh + uh < oldh) ? 1u : 0u);
    if (e) atomicAdd(&ext[slot], e);
}
#define QG_WIDE_FROM 33554432.0     /* 2^25 value units: contributions of this size and more take bin_add96_wide (the warp-aggregated
                                       sum of 32 smaller ones still fits 64 bits) */
__device__ __forceinline__ double bin_value96(unsigned lo, unsigned hi, unsigned ext) {
    return (double)(long long)(((unsigned long long)ext << 32) | (unsigned long long)hi) + (double)lo * (1.0 / 4294967296.0);
}
#define QG_SCALE 4294967296.0
#ifndef QG_RUN
#define QG_RUN 4        /* consecutive 64-quartet tiles a warp takes before the next warp's run begins */
#endif
// The same pass for SMALL tables (fewer than 1024 class-A tiles of 64 quartets): rows dealt to the warps one by one,
// plain loads, per-block gradient rows added by the last block.  A small table is all latency: no ring to set up, no
// atomics to drain, the shortest dependent chain per warp (k_eval_grad below is the streaming variant for large tables).
template <bool WANT_H>
__global__ void __launch_bounds__(256, 3)
k_eval_grad_small(SnaqTables T, const uint32_t *__restrict__ off, const uint2 *__restrict__ prog2, const uint8_t *__restrict__ qhdr,
            const double2 *__restrict__ WA, const double4 *__restrict__ W, uint32_t nA1, uint32_t nA2, uint32_t a2base, uint32_t nQ,
            const double *__restrict__ X, const double *__restrict__ xconst, int nh, int nt, double *__restrict__ partial,
            double *__restrict__ out, const double *__restrict__ negw3, unsigned int *__restrict__ ticket,
            double *__restrict__ gpart, double *__restrict__ gout, uint32_t *__restrict__ status, uint32_t *__restrict__ status_out,
            const int use_xarg, const double *__restrict__ pen, const __grid_constant__ XArg xarg) {
    constexpr int BLOCK = 256;
    extern __shared__ __align__(128) unsigned char smem[];
    double *mt = reinterpret_cast<double *>(smem);                        // 48 doubles, 128-byte aligned
    double *red = reinterpret_cast<double *>(smem + 384);                 // [8] warp sums
    uint2 *span = reinterpret_cast<uint2 *>(smem + 512);                  // [8 warps][QD_SPAN] staged programs
    double *xs = reinterpret_cast<double *>(smem + 512 + 8 * QD_SPAN * 8);
    const int n_full = T.n_full;
    unsigned int *bins_lo = reinterpret_cast<unsigned int *>(xs + n_full);   // [n_full] low words
    unsigned int *bins_hi = bins_lo + n_full;                                // [n_full] middle words
    unsigned int *bins_ex = bins_hi + n_full;                                // [n_full] top words (see bin_add96)
    unsigned int *hb_lo = bins_ex + n_full, *hb_hi = hb_lo + n_full, *hb_ex = hb_hi + n_full;   // WANT_H: curvature bins (same format)
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    unsigned st = 0;
    const SnaqMath M = load_math_tables(mt, (int)tid);
    double *xraw = reinterpret_cast<double *>(bins_lo + (((WANT_H ? 6 : 3) * n_full + 1) & ~1));     // [nparams] the candidate itself
    for (int row = (int)tid; row < T.nparams; row += BLOCK) {
        const double v = use_xarg ? xarg.v[row] : X[row];
        st |= range_check(v, (row >= nh && row < nh + nt) ? 1.0e300 : 1.0);
        xraw[row] = v;
    }
    __syncthreads();
    for (int row = (int)tid; row < n_full; row += BLOCK) {
        xs[row] = snaq_expand_row(T, M, xraw, xconst, row);
        bins_lo[row] = 0u;
        bins_hi[row] = 0u;
        bins_ex[row] = 0u;
        if (WANT_H) { hb_lo[row] = 0u; hb_hi[row] = 0u; hb_ex[row] = 0u; }
    }
    __syncthreads();
    const unsigned gwarp = blockIdx.x * (BLOCK / 32) + warp, nwarps = gridDim.x * (BLOCK / 32);
    const unsigned nul = (unsigned)T.n_xe;
    double acc = 0.0;
    auto addq = [&](unsigned slot, long long q) {
        if (slot != nul) bin_add96(bins_lo, bins_hi, bins_ex, slot, q);
    };
    auto add = [&](unsigned slot, double v) {
        if (fabs(v) < QG_WIDE_FROM) addq(slot, __double2ll_rn(v * QG_SCALE));
        else if (slot != nul) bin_add96_wide(bins_lo, bins_hi, bins_ex, slot, v);
    };
    auto addhq = [&](unsigned slot, long long q) {
        if (WANT_H && slot != nul) bin_add96(hb_lo, hb_hi, hb_ex, slot, q);
    };
    auto addh = [&](unsigned slot, double v) {
        if (!WANT_H) return;
        if (fabs(v) < QG_WIDE_FROM) addhq(slot, __double2ll_rn(v * QG_SCALE));
        else if (slot != nul) bin_add96_wide(hb_lo, hb_hi, hb_ex, slot, v);
    };
    auto tree = [&](double t, const double2 &w, long long *gq, long long *hq) {   // additive quartet: value, dv/dt, -d2v/dt2
        const double t1 = snaq_clamp10(t);
        const double a = 2.0 / 3.0 * snaq_exp(-t1, M);
        const double major = 1.0 - a;
        const double r = w.x * a / major;
        const bool in = t <= 10.0;
        *gq = in ? __double2ll_rn((r - w.y) * QG_SCALE) : 0ll;
        if (WANT_H) *hq = in ? __double2ll_rn(r / major * QG_SCALE) : 0ll;
        return fma(w.x, snaq_log_pos(major, M), -w.y * (t1 + SNAQ_K_LOG3));
    };
    // A warp holds 32 CONSECUTIVE entries of the table, and class A is sorted by program: most of the time the 32 lanes
    // read the SAME slots.  wadd: when a slot is warp-uniform the 32 values are summed with integer warp reductions (three
    // 21-bit limbs, REDUX) and lane 0 issues ONE atomic; otherwise every lane adds its own.
    auto warp_sum = [&](long long q) {
        const unsigned l0 = (unsigned)(q & 0x1FFFFFll), l1 = (unsigned)((q >> 21) & 0x1FFFFFll);
        const int l2 = (int)(q >> 42);
        const unsigned s0 = __reduce_add_sync(0xffffffffu, l0), s1 = __reduce_add_sync(0xffffffffu, l1);
        const int s2 = __reduce_add_sync(0xffffffffu, l2);
        return (long long)s0 + ((long long)s1 << 21) + ((long long)s2 << 42);
    };
    auto wadd = [&](unsigned slot, long long q, long long hq) {          // called by all 32 lanes
        const unsigned s0 = __shfl_sync(0xffffffffu, slot, 0);
        if (__all_sync(0xffffffffu, slot == s0)) {
            if (s0 == nul) return;
            const long long tq = warp_sum(q);
            long long th = 0;
            if (WANT_H) th = warp_sum(hq);
            if (lane == 0) { if (tq != 0) addq(s0, tq); if (WANT_H && th != 0) addhq(s0, th); }
        } else {
            if (q != 0) addq(slot, q);
            if (WANT_H && hq != 0) addhq(slot, hq);
        }
    };
    // ---- region A1: one chunk (two signed pairs) ---------------------------------------------------------
#pragma unroll 1
    for (unsigned ib = gwarp * 32u; ib < nA1; ib += 2u * nwarps * 32u) {
        double2 w[2];
        uint2 c[2];
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const unsigned i = ib + (unsigned)k * nwarps * 32u + lane;
            const bool ok = i < nA1;
            w[k] = ok ? __ldg(WA + i) : make_double2(0.0, 0.0);
            c[k] = ok ? __ldg(prog2 + i) : make_uint2(nul * 0x10001u, nul * 0x10001u);
        }
#pragma unroll
        for (int k = 0; k < 2; k++) {
            if (ib + (unsigned)k * nwarps * 32u >= nA1) break;           // warp-uniform
            const unsigned a = c[k].x & 0xffffu, b = c[k].x >> 16, cc = c[k].y & 0xffffu, d = c[k].y >> 16;
            long long gq, hq = 0;
            acc += tree((xs[a] - xs[b]) + (xs[cc] - xs[d]), w[k], &gq, &hq);
            // the whole program is usually the same in all 32 lanes (the class is sorted by program): one reduction, then
            // lane 0 adds the warp's total to the four slots
            const unsigned px = __shfl_sync(0xffffffffu, c[k].x, 0), py = __shfl_sync(0xffffffffu, c[k].y, 0);
            if (__all_sync(0xffffffffu, c[k].x == px && c[k].y == py)) {
                const long long tq = warp_sum(gq);
                long long th = 0;
                if (WANT_H) th = warp_sum(hq);
                if (lane == 0) {
                    if (tq != 0) { addq(a, tq); addq(b, -tq); addq(cc, tq); addq(d, -tq); }
                    if (WANT_H && th != 0) { addhq(a, th); addhq(b, -th); addhq(cc, th); addhq(d, -th); }
                }
            } else { wadd(a, gq, hq); wadd(b, -gq, -hq); wadd(cc, gq, hq); wadd(d, -gq, -hq); }
        }
    }
    // ---- region A2: two chunks -----------------------------------------------------------------------------
#pragma unroll 1
    for (unsigned jb = gwarp * 32u; jb < nA2; jb += nwarps * 32u) {
        const unsigned j = jb + lane;
        const bool ok = j < nA2;
        const double2 w = ok ? __ldg(WA + nA1 + j) : make_double2(0.0, 0.0);
        const unsigned np2 = nul * 0x10001u;
        const uint4 c = ok ? __ldg(reinterpret_cast<const uint4 *>(prog2 + a2base) + j) : make_uint4(np2, np2, np2, np2);
        const unsigned s[8] = {c.x & 0xffffu, c.x >> 16, c.y & 0xffffu, c.y >> 16, c.z & 0xffffu, c.z >> 16, c.w & 0xffffu, c.w >> 16};
        double ta = xs[s[0]] - xs[s[1]], tb = xs[s[2]] - xs[s[3]];
        ta += xs[s[4]] - xs[s[5]];
        tb += xs[s[6]] - xs[s[7]];
        long long gq, hq = 0;
        acc += tree(ta + tb, w, &gq, &hq);
        const unsigned p0 = __shfl_sync(0xffffffffu, c.x, 0), p1 = __shfl_sync(0xffffffffu, c.y, 0);
        const unsigned p2 = __shfl_sync(0xffffffffu, c.z, 0), p3 = __shfl_sync(0xffffffffu, c.w, 0);
        if (__all_sync(0xffffffffu, c.x == p0 && c.y == p1 && c.z == p2 && c.w == p3)) {
            const long long tq = warp_sum(gq);
            long long th = 0;
            if (WANT_H) th = warp_sum(hq);
            if (lane == 0) {
#pragma unroll
                for (int k = 0; k < 8; k += 2) {
                    if (tq != 0) { addq(s[k], tq); addq(s[k + 1], -tq); }
                    if (WANT_H && th != 0) { addhq(s[k], th); addhq(s[k + 1], -th); }
                }
            }
        } else {
#pragma unroll
            for (int k = 0; k < 8; k += 2) { wadd(s[k], gq, hq); wadd(s[k + 1], -gq, -hq); }
        }
    }
    // ---- region G: everything else through the generic evaluator ------------------------------------------
    {
        uint2 *wspan = span + warp * QD_SPAN;
#pragma unroll 1
        for (unsigned ib = nA1 + nA2 + gwarp * 32u; ib < nQ; ib += nwarps * 32u) {
            const unsigned i = ib + lane;
            const bool active = i < nQ;
            const unsigned ilast = min(ib + 32u, nQ);
            const uint32_t o0 = active ? __ldg(off + i) : 0u, o1 = active ? __ldg(off + i + 1) : 0u;
            const unsigned hdr = active ? (unsigned)qhdr[i] : 0u;
            const double4 w4 = active ? W[i] : make_double4(0, 0, 0, 0);
            const uint32_t s0 = __shfl_sync(0xffffffffu, o0, 0);
            const uint32_t s1 = __ldg(off + ilast);
            const bool staged = (s1 - s0) <= (uint32_t)QD_SPAN;
            __syncwarp();
            if (staged) for (uint32_t c = s0 + lane; c < s1; c += 32u) wspan[c - s0] = __ldg(prog2 + c);
            __syncwarp();
            if (active) {
                const double w[4] = {w4.x, w4.y, w4.z, w4.w};
                const uint16_t *pc = reinterpret_cast<const uint16_t *>(staged ? wspan + (o0 - s0) : prog2 + o0);
                const XRow xf{xs};
                acc += snaq_grad_quartet(hdr, pc, (int)(o1 - o0) * 4, xf, M, w, nul, add, addh, st, pen ? pen + 4 * (size_t)i : nullptr);
            }
        }
    }
    if (st) atomicOr(status, st);
    // ---- epilogue ---------------------------------------------------------------------------------------------
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if (lane == 0) red[warp] = acc;
    __syncthreads();                                                      // also: every atomic on bins is done
    if (tid == 0) {
        double ssum = 0.0;
#pragma unroll
        for (int k = 0; k < BLOCK / 32; k++) ssum += red[k];
        partial[blockIdx.x] = ssum;
    }
    const int nrows = WANT_H ? 2 * n_full : n_full;                       // the curvature bins follow the gradient bins
    for (int row = (int)tid; row < nrows; row += BLOCK)
        gpart[(size_t)blockIdx.x * nrows + row] = row < n_full ? bin_value96(bins_lo[row], bins_hi[row], bins_ex[row])
                                                               : bin_value96(hb_lo[row - n_full], hb_hi[row - n_full], hb_ex[row - n_full]);
    __threadfence();
    __shared__ unsigned int s_last;
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1) ? 1u : 0u;
    __syncthreads();
    if (s_last) {
        __threadfence();
        double sp = 0.0;
        for (unsigned b = tid; b < gridDim.x; b += BLOCK) sp += __ldcg(partial + b);
        for (int o = 16; o > 0; o >>= 1) sp += __shfl_down_sync(0xffffffffu, sp, o);
        if (lane == 0) red[warp] = sp;
        __syncthreads();
        if (tid == 0) {
            double tot = 0.0;
#pragma unroll
            for (int k = 0; k < BLOCK / 32; k++) tot += red[k];
            out[0] = -tot - __ldcg(negw3);
        }
        for (int row = (int)tid; row < nrows; row += BLOCK) {
            double s = 0.0;
            for (unsigned b = 0; b < gridDim.x; b++) s += __ldcg(gpart + (size_t)b * nrows + row);
            gout[row] = row < n_full ? -s : s;                            // f = -(sum of v); the curvature is already that of f
        }
        __threadfence_system();                          // every thread's rows are in host memory before the flag
        __syncthreads();
        if (tid == 0) {
            *ticket = 0u;
            if (status_out) *reinterpret_cast<volatile uint32_t *>(status_out) = *reinterpret_cast<volatile uint32_t *>(status) | SNAQ_S_DONE;
        }
    }
}

template <bool WANT_H>
__global__ void __launch_bounds__(256, 3)
k_eval_grad(SnaqTables T, const uint32_t *__restrict__ off, const uint2 *__restrict__ prog2, const uint8_t *__restrict__ qhdr,
            const double2 *__restrict__ WA, const double4 *__restrict__ W, uint32_t nA1, uint32_t nA2, uint32_t a2base, uint32_t nQ,
            const double *__restrict__ X, const double *__restrict__ xconst, int nh, int nt, double *__restrict__ partial,
            double *__restrict__ out, const double *__restrict__ negw3, unsigned int *__restrict__ ticket,
            unsigned long long *__restrict__ gbins, double *__restrict__ gout, uint32_t *__restrict__ status, uint32_t *__restrict__ status_out,
            const int use_xarg /* bit 0: x is in xarg; bit 1: no warp aggregation in the generic path */,
            const double *__restrict__ pen,
            unsigned long long *__restrict__ trace, const __grid_constant__ XArg xarg) {
    constexpr int BLOCK = 256;
    extern __shared__ __align__(128) unsigned char smem[];
    // optional phase timeline (SNAQ_TRACE=1): entry / prologue / A1 / A2 / G / partials written / exit
    auto stamp = [&](int k) {
        if (trace && threadIdx.x == 0) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            trace[(size_t)blockIdx.x * 8 + k] = t;
        }
    };
    stamp(0);
    double *mt = reinterpret_cast<double *>(smem);                        // 48 doubles, 128-byte aligned
    double *red = reinterpret_cast<double *>(smem + 384);                 // [8] warp sums (64 B)
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + 448);            // [8 warps][QD_NST] mbarriers (192 B)
    uint2 *span = reinterpret_cast<uint2 *>(smem + 640);                  // [8 warps][QD_SPAN] staged programs
    unsigned char *ring = smem + 640 + 8 * QD_SPAN * 8;                   // [8 warps][QD_NST][QD_STAGE_BYTES] class-A tiles
    double *xs = reinterpret_cast<double *>(ring + 8 * QD_NST * QD_STAGE_BYTES);
    const int n_full = T.n_full;
    unsigned int *bins_lo = reinterpret_cast<unsigned int *>(xs + n_full);   // [n_full] low words
    unsigned int *bins_hi = bins_lo + n_full;                                // [n_full] middle words
    unsigned int *bins_ex = bins_hi + n_full;                                // [n_full] top words (see bin_add96)
    unsigned int *hb_lo = bins_ex + n_full, *hb_hi = hb_lo + n_full, *hb_ex = hb_hi + n_full;   // WANT_H: curvature bins (same format)
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    unsigned st = 0;
    // class A (one- and two-chunk additive programs) is streamed through a warp-private ring of TMA bulk copies; every
    // warp walks runs of consecutive 64-quartet tiles of the class (the class is sorted by program)
    const unsigned gwarp = blockIdx.x * (BLOCK / 32) + warp, nwarps = gridDim.x * (BLOCK / 32);
    const unsigned tA1 = (nA1 + QD_TQ - 1) / QD_TQ, tA2 = (nA2 + QD_TQ - 1) / QD_TQ, tA = tA1 + tA2;
    // the k-th tile of this warp: runs of QG_RUN consecutive tiles, the runs dealt round-robin to the warps (consecutive
    // tiles share programs, which is what the memo below lives on; dealing short runs evens out the cost per warp: a tile
    // full of short program runs costs ten times a tile inside one long run)
    // (small classes: shorter runs, so that every warp has a tile before any warp has two)
    const unsigned run = max(1u, min((unsigned)QG_RUN, tA / nwarps));
    auto tile_of = [&](unsigned k) { return (gwarp + (k / run) * nwarps) * run + k % run; };
    uint64_t *wbar = bars + warp * QD_NST;
    unsigned char *wring = ring + (size_t)warp * QD_NST * QD_STAGE_BYTES;
    auto issue = [&](unsigned t, unsigned sg) {                           // lane 0 only: weight pairs, then program chunks
        unsigned char *dst = wring + sg * QD_STAGE_BYTES;
        if (t < tA1) {
            const unsigned i0 = t * QD_TQ, n = min((unsigned)QD_TQ, nA1 - i0);
            const unsigned pb = ((n + 1u) & ~1u) * 8u;
            mbar_expect_tx(&wbar[sg], n * 16u + pb);
            tma_load_1d(dst, WA + i0, n * 16u, &wbar[sg]);
            tma_load_1d(dst + QD_TQ * 16, prog2 + i0, pb, &wbar[sg]);
        } else {
            const unsigned j0 = (t - tA1) * QD_TQ, n = min((unsigned)QD_TQ, nA2 - j0);
            mbar_expect_tx(&wbar[sg], n * 32u);
            tma_load_1d(dst, WA + nA1 + j0, n * 16u, &wbar[sg]);
            tma_load_1d(dst + QD_TQ * 16, prog2 + a2base + 2u * j0, n * 16u, &wbar[sg]);
        }
    };
    if (lane == 0) {
        for (int k = 0; k < QD_NST; k++) mbar_init(&wbar[k], 1);
        for (unsigned k = 0; k < QD_NST; k++)
            if (tile_of(k) < tA) issue(tile_of(k), k);
    }
    __syncwarp();
    const SnaqMath M = load_math_tables(mt, (int)tid);
    double *xraw = reinterpret_cast<double *>(bins_lo + (((WANT_H ? 6 : 3) * n_full + 1) & ~1));     // [nparams] the candidate itself
    for (int row = (int)tid; row < T.nparams; row += BLOCK) {
        const double v = (use_xarg & 1) ? xarg.v[row] : X[row];
        st |= range_check(v, (row >= nh && row < nh + nt) ? 1.0e300 : 1.0);
        xraw[row] = v;
    }
    __syncthreads();
    for (int row = (int)tid; row < n_full; row += BLOCK) {
        xs[row] = snaq_expand_row(T, M, xraw, xconst, row);
        bins_lo[row] = 0u;
        bins_hi[row] = 0u;
        bins_ex[row] = 0u;
        if (WANT_H) { hb_lo[row] = 0u; hb_hi[row] = 0u; hb_ex[row] = 0u; }
    }
    __syncthreads();
    stamp(1);
    const unsigned nul = (unsigned)T.n_xe;
    double acc = 0.0;
    auto addq = [&](unsigned slot, long long q) {
        if (slot != nul) bin_add96(bins_lo, bins_hi, bins_ex, slot, q);
    };
    // generic path: the lanes of a warp mostly run the SAME program (the classes are sorted by term and program), so they
    // add to the same slot at the same instruction; 32 atomics on one shared-memory word are served one after the other.
    // Lanes that are converged here and hold the same slot sum their fixed-point values with integer warp reductions
    // (exact, so the bins stay order independent) and the group's first lane issues the one atomic.
    const bool agg_on = (use_xarg & 2) == 0;                  // off for deduplicated (neighbouring programs all differ) and small tables
    auto agg = [&](unsigned slot, long long q, bool curv) {
        // (only the all-equal case is aggregated: shuffle + vote; a general match.any costs more than the conflicts)
        unsigned peers = 1u << lane;
        if (agg_on) {
            const unsigned act = __activemask();
            const unsigned s0 = __shfl_sync(act, slot, __ffs(act) - 1);
            if (__all_sync(act, slot == s0)) peers = act;
        }
        if (peers != (1u << lane)) {
            const unsigned l0 = (unsigned)(q & 0x1FFFFFll), l1 = (unsigned)((q >> 21) & 0x1FFFFFll);
            const int l2 = (int)(q >> 42);
            const unsigned s0 = __reduce_add_sync(peers, l0), s1 = __reduce_add_sync(peers, l1);
            const int s2 = __reduce_add_sync(peers, l2);
            if (lane != (unsigned)(__ffs(peers) - 1)) return;
            q = (long long)s0 + ((long long)s1 << 21) + ((long long)s2 << 42);
        }
        if (q == 0) return;
        if (curv) bin_add96(hb_lo, hb_hi, hb_ex, slot, q);
        else bin_add96(bins_lo, bins_hi, bins_ex, slot, q);
    };
    auto add = [&](unsigned slot, double v) {
        if (slot == nul) return;
        if (fabs(v) < QG_WIDE_FROM) agg(slot, __double2ll_rn(v * QG_SCALE), false);
        else bin_add96_wide(bins_lo, bins_hi, bins_ex, slot, v);
    };
    auto addhq = [&](unsigned slot, long long q) {
        if (WANT_H && slot != nul) bin_add96(hb_lo, hb_hi, hb_ex, slot, q);
    };
    auto addh = [&](unsigned slot, double v) {
        if (!WANT_H || slot == nul) return;
        if (fabs(v) < QG_WIDE_FROM) agg(slot, __double2ll_rn(v * QG_SCALE), true);
        else bin_add96_wide(hb_lo, hb_hi, hb_ex, slot, v);
    };
    // additive quartet: everything that depends on the program only (t -> log of the major CF, t1 + log 3, the two
    // derivative factors); the quartet's own weights enter linearly
    struct TreeConsts { double L, T, r, h; bool in; };
    auto tree_consts = [&](double t) {
        TreeConsts k;
        const double t1 = snaq_clamp10(t);
        const double a = 2.0 / 3.0 * snaq_exp(-t1, M);
        const double major = 1.0 - a;
        k.L = snaq_log_pos(major, M);
        k.T = t1 + SNAQ_K_LOG3;
        k.r = a / major;
        k.h = k.r / major;
        k.in = t <= 10.0;
        return k;
    };
    // A warp holds 32 CONSECUTIVE entries of the table, and class A is sorted by program: most of the time the 32 lanes
    // read the SAME slots.  wadd: when a slot is warp-uniform the 32 values are summed with integer warp reductions (three
    // 21-bit limbs, REDUX) and lane 0 issues ONE atomic; otherwise every lane adds its own.
    auto warp_sum = [&](long long q) {
        const unsigned l0 = (unsigned)(q & 0x1FFFFFll), l1 = (unsigned)((q >> 21) & 0x1FFFFFll);
        const int l2 = (int)(q >> 42);
        const unsigned s0 = __reduce_add_sync(0xffffffffu, l0), s1 = __reduce_add_sync(0xffffffffu, l1);
        const int s2 = __reduce_add_sync(0xffffffffu, l2);
        return (long long)s0 + ((long long)s1 << 21) + ((long long)s2 << 42);
    };
    auto wadd = [&](unsigned slot, long long q, long long hq) {          // called by all 32 lanes
        const unsigned s0 = __shfl_sync(0xffffffffu, slot, 0);
        if (__all_sync(0xffffffffu, slot == s0)) {
            if (s0 == nul) return;
            const long long tq = warp_sum(q);
            long long th = 0;
            if (WANT_H) th = warp_sum(hq);
            if (lane == 0) { if (tq != 0) addq(s0, tq); if (WANT_H && th != 0) addhq(s0, th); }
        } else {
            if (q != 0) addq(slot, q);
            if (WANT_H && hq != 0) addhq(slot, hq);
        }
    };
    // ---- region A: the warp's tiles, 32 consecutive entries (a row) at a time --------------------------------
    // A row usually holds ONE program, and so do the rows after it: the program's constants are computed once (memo) and
    // every lane keeps its quartets' fixed-point derivative sums in registers; the warp reduction and the atomics on the
    // block's bins happen only when the program changes.  The sums are integers: the bins do not depend on where the
    // flushes fall.  A one-chunk program is handled as a two-chunk program whose second chunk reads the null slot.
    {
        const unsigned nullpair = nul * 0x10001u;
        TreeConsts mk{0.0, 0.0, 0.0, 0.0, false};
        long long gacc = 0, hacc = 0;
        bool m_on = false;
        uint4 mp = make_uint4(0u, 0u, 0u, 0u);
        auto slots = [&](const uint4 &c, unsigned *sl) {
            sl[0] = c.x & 0xffffu; sl[1] = c.x >> 16; sl[2] = c.y & 0xffffu; sl[3] = c.y >> 16;
            sl[4] = c.z & 0xffffu; sl[5] = c.z >> 16; sl[6] = c.w & 0xffffu; sl[7] = c.w >> 16;
        };
        auto tsum = [&](const unsigned *sl) {
            double ta = xs[sl[0]] - xs[sl[1]], tb2 = xs[sl[2]] - xs[sl[3]];
            ta += xs[sl[4]] - xs[sl[5]];
            tb2 += xs[sl[6]] - xs[sl[7]];
            return ta + tb2;
        };
        auto flush = [&]() {
            if (!m_on) return;
            const long long tq = warp_sum(gacc);
            long long th = 0;
            if (WANT_H) th = warp_sum(hacc);
            if (lane == 0) {
                unsigned sl[8];
                slots(mp, sl);
#pragma unroll
                for (int k = 0; k < 8; k += 2) {
                    if (tq != 0) { addq(sl[k], tq); addq(sl[k + 1], -tq); }
                    if (WANT_H && th != 0) { addhq(sl[k], th); addhq(sl[k + 1], -th); }
                }
            }
            gacc = 0; hacc = 0;
        };
        unsigned sg = 0, phase = 0;
#pragma unroll 1
        for (unsigned kt = 0;; kt++) {
            const unsigned t = tile_of(kt);
            if (t >= tA) break;
            mbar_wait(&wbar[sg], phase);
            const unsigned char *stg = wring + sg * QD_STAGE_BYTES;
            const bool one = t < tA1;
            const unsigned q0 = one ? t * QD_TQ : (t - tA1) * QD_TQ;     // first entry of the tile inside its sub-class
            const unsigned n = one ? min((unsigned)QD_TQ, nA1 - q0) : min((unsigned)QD_TQ, nA2 - q0);
            const double2 *sw = reinterpret_cast<const double2 *>(stg);    // the tile's weight pairs, then its programs
            const uint2 *sp2 = reinterpret_cast<const uint2 *>(stg + QD_TQ * 16);
            const uint4 *sp4 = reinterpret_cast<const uint4 *>(stg + QD_TQ * 16);
#pragma unroll 1
            for (unsigned r0 = 0; r0 < n; r0 += 32u) {
                const unsigned idx = r0 + lane;
                const bool ok = idx < n;
                const double2 w = ok ? sw[idx] : make_double2(0.0, 0.0);
                uint4 c = make_uint4(0u, 0u, nullpair, nullpair);
                if (ok) {
                    if (one) { const uint2 c2 = sp2[idx]; c.x = c2.x; c.y = c2.y; }
                    else c = sp4[idx];
                }
                uint4 pz;                                                // the row's first program
                pz.x = __shfl_sync(0xffffffffu, c.x, 0); pz.y = __shfl_sync(0xffffffffu, c.y, 0);
                pz.z = __shfl_sync(0xffffffffu, c.z, 0); pz.w = __shfl_sync(0xffffffffu, c.w, 0);
                if (!ok) c = pz;                                         // lanes past the end: zero weights, add nothing
                if (__all_sync(0xffffffffu, c.x == pz.x && c.y == pz.y && c.z == pz.z && c.w == pz.w)) {
                    if (!(m_on && pz.x == mp.x && pz.y == mp.y && pz.z == mp.z && pz.w == mp.w)) {
                        flush();
                        mp = pz; m_on = true;
                        unsigned sl[8];
                        slots(pz, sl);
                        mk = tree_consts(tsum(sl));
                    }
                    acc += fma(w.x, mk.L, -w.y * mk.T);
                    if (mk.in) {
                        gacc += __double2ll_rn(fma(mk.r, w.x, -w.y) * QG_SCALE);
                        if (WANT_H) hacc += __double2ll_rn(mk.h * w.x * QG_SCALE);
                    }
                } else {
                    flush();
                    m_on = false;
                    unsigned sl[8];
                    slots(c, sl);
                    const TreeConsts k = tree_consts(tsum(sl));
                    acc += fma(w.x, k.L, -w.y * k.T);
                    const long long gq = k.in ? __double2ll_rn(fma(k.r, w.x, -w.y) * QG_SCALE) : 0ll;
                    const long long hq = (WANT_H && k.in) ? __double2ll_rn(k.h * w.x * QG_SCALE) : 0ll;
#pragma unroll
                    for (int q = 0; q < 8; q += 2) { wadd(sl[q], gq, hq); wadd(sl[q + 1], -gq, -hq); }
                }
            }
            __syncwarp();                                                  // every lane is done with stage sg
            if (lane == 0 && tile_of(kt + QD_NST) < tA) issue(tile_of(kt + QD_NST), sg);
            if (++sg == QD_NST) { sg = 0; phase ^= 1u; }
        }
        flush();
    }
    stamp(2);
    stamp(3);
    // ---- region G: everything else through the generic evaluator ------------------------------------------
    {
        uint2 *wspan = span + warp * QD_SPAN;
#pragma unroll 1
        for (unsigned ib = nA1 + nA2 + gwarp * 32u; ib < nQ; ib += nwarps * 32u) {
            const unsigned i = ib + lane;
            const bool active = i < nQ;
            const unsigned ilast = min(ib + 32u, nQ);
            const uint32_t o0 = active ? __ldg(off + i) : 0u, o1 = active ? __ldg(off + i + 1) : 0u;
            const unsigned hdr = active ? (unsigned)qhdr[i] : 0u;
            const double4 w4 = active ? W[i] : make_double4(0, 0, 0, 0);
            const uint32_t s0 = __shfl_sync(0xffffffffu, o0, 0);
            const uint32_t s1 = __ldg(off + ilast);
            const bool staged = (s1 - s0) <= (uint32_t)QD_SPAN;
            __syncwarp();
            if (staged) for (uint32_t c = s0 + lane; c < s1; c += 32u) wspan[c - s0] = __ldg(prog2 + c);
            __syncwarp();
            if (active) {
                const double w[4] = {w4.x, w4.y, w4.z, w4.w};
                const uint16_t *pc = reinterpret_cast<const uint16_t *>(staged ? wspan + (o0 - s0) : prog2 + o0);
                const XRow xf{xs};
                acc += snaq_grad_quartet(hdr, pc, (int)(o1 - o0) * 4, xf, M, w, nul, add, addh, st, pen ? pen + 4 * (size_t)i : nullptr);
            }
        }
    }
    if (st) atomicOr(status, st);
    stamp(4);
    // ---- epilogue ---------------------------------------------------------------------------------------------
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if (lane == 0) red[warp] = acc;
    __syncthreads();                                                      // also: every atomic on bins is done
    stamp(5);
    if (tid == 0) {
        double ssum = 0.0;
#pragma unroll
        for (int k = 0; k < BLOCK / 32; k++) ssum += red[k];
        partial[blockIdx.x] = ssum;
    }
    const int nrows = WANT_H ? 2 * n_full : n_full;                       // the curvature bins follow the gradient bins
    // the block's bins go to the grid's bins with 64-bit integer reductions (order independent, so bit-reproducible):
    // gbins[row] takes the fraction words (2^32 blocks of headroom), gbins[nrows + row] the blocks' signed 64-bit integer
    // parts (top and middle word of the 96-bit bins); the last block only converts (no serial pass over per-block rows)
    for (int row = (int)tid; row < nrows; row += BLOCK) {
        const bool cv = row >= n_full;
        const int r = cv ? row - n_full : row;
        const unsigned lo = cv ? hb_lo[r] : bins_lo[r], hi = cv ? hb_hi[r] : bins_hi[r], ex = cv ? hb_ex[r] : bins_ex[r];
        if (lo) atomicAdd(gbins + row, (unsigned long long)lo);
        if (hi | ex) atomicAdd(gbins + nrows + row, ((unsigned long long)ex << 32) | (unsigned long long)hi);
    }
    stamp(7);
    __threadfence();
    __shared__ unsigned int s_last;
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1) ? 1u : 0u;
    __syncthreads();
    if (s_last) {
        __threadfence();
        double sp = 0.0;
        for (unsigned b = tid; b < gridDim.x; b += BLOCK) sp += __ldcg(partial + b);
        for (int o = 16; o > 0; o >>= 1) sp += __shfl_down_sync(0xffffffffu, sp, o);
        if (lane == 0) red[warp] = sp;
        __syncthreads();
        if (tid == 0) {
            double tot = 0.0;
#pragma unroll
            for (int k = 0; k < BLOCK / 32; k++) tot += red[k];
            out[0] = -tot - __ldcg(negw3);
        }
        for (int row = (int)tid; row < nrows; row += BLOCK) {
            const unsigned long long fr = __ldcg(gbins + row);
            const long long in = (long long)__ldcg(gbins + nrows + row);
            gbins[row] = 0ull;                                            // ready for the next launch
            gbins[nrows + row] = 0ull;
            const double s = (double)in + (double)fr * (1.0 / QG_SCALE);
            gout[row] = row < n_full ? -s : s;                            // f = -(sum of v); the curvature is already that of f
        }
        __threadfence_system();                          // every thread's rows are in host memory before the flag
        __syncthreads();
        if (tid == 0) {
            *ticket = 0u;
            if (status_out) *reinterpret_cast<volatile uint32_t *>(status_out) = *reinterpret_cast<volatile uint32_t *>(status) | SNAQ_S_DONE;
        }
    }
    stamp(6);
}

// Weighted sampling of quartets (step 1 of sampleEdgeQuartetWeighted, src/moves_snaq.jl:1443-1449): inverse-CDF walk over
// the deltaCF weights.  k_chunk_sums: one block per chunk of SQ_CHUNK weights, every thread adds its 16 consecutive weights
// serially, then a fixed-order tree: the same bits on every run.  k_sample: thread 0 turns the chunk totals into running
// totals (serial, in shared memory), then one warp per draw finds the chunk by bisection and walks it 32 entries at a time.
#define SQ_CHUNK 4096
__global__ void __launch_bounds__(256) k_chunk_sums(const double *__restrict__ w, const int64_t nq, double *__restrict__ sums) {
    __shared__ double red[8];
    const int64_t base = (int64_t)blockIdx.x * SQ_CHUNK + (int64_t)threadIdx.x * (SQ_CHUNK / 256);
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < SQ_CHUNK / 256; k++) s += (base + k < nq) ? w[base + k] : 0.0;
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
#pragma unroll
        for (int k = 0; k < 8; k++) t += red[k];
        sums[blockIdx.x] = t;
    }
}
__global__ void __launch_bounds__(256) k_sample(const double *__restrict__ w, const int64_t nq, double *__restrict__ cum,
                                                const int nchunks, const int in_smem, const double *__restrict__ u, const int n,
                                                long long *__restrict__ idx) {
    extern __shared__ double s_cum[];
    double *c = in_smem ? s_cum : cum;                       // running chunk totals: shared memory when they fit
    if (in_smem) for (int k = threadIdx.x; k < nchunks; k += 256) s_cum[k] = cum[k];
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int k = 0; k < nchunks; k++) { t += c[k]; c[k] = t; }          // serial, as the reference's walk
    }
    __syncthreads();
    const double total = c[nchunks - 1];
    const unsigned lane = threadIdx.x & 31u;
    for (int k = threadIdx.x >> 5; k < n; k += 8) {          // one warp per draw
        const double t = u[k] * total;                       // t = rand() * sum(wv)
        int lo = 0, hi = nchunks - 1;                        // first chunk whose running total reaches t (the last one if none)
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (c[mid] >= t) hi = mid; else lo = mid + 1;
        }
        double cw = lo > 0 ? c[lo - 1] : 0.0;
        const int64_t i0 = (int64_t)lo * SQ_CHUNK, i1 = min(i0 + (int64_t)SQ_CHUNK, nq);
        // `while cw < t && i < n` of the reference's walk, 32 entries at a time: inclusive scan of the 32 weights in a
        // fixed order, first lane whose running sum reaches t
        long long found = -1;
        for (int64_t b = i0; b < i1 && found < 0; b += 32) {
            const int64_t i = b + lane;
            double v = i < i1 ? w[i] : 0.0;
            for (int o = 1; o < 32; o <<= 1) {
                const double y = __shfl_up_sync(0xffffffffu, v, o);
                if (lane >= (unsigned)o) v += y;
            }
            const unsigned hit = __ballot_sync(0xffffffffu, i < i1 && cw + v >= t);
            if (hit) found = b + (__ffs(hit) - 1);
            cw += __shfl_sync(0xffffffffu, v, 31);
        }
        if (found < 0) {
            // rounding between the chunk total and the walk can leave cw < t at the end of a chunk: the draw belongs to the
            // first quartet of positive weight after it (the last quartet if there is none, as `i < n` ends the walk)
            found = i1 - 1;
            if (i1 < nq) {
                if (lane == 0) { int64_t i = i1; while (i < nq - 1 && !(w[i] > 0.0)) i++; found = i; }
                found = __shfl_sync(0xffffffffu, found, 0);
            }
        }
        if (lane == 0) idx[k] = found;
    }
}

// updateBL! (src/readquartetdata.jl:838-861): per internal edge of a TREE, the number of data quartets whose internal
// path is exactly that edge (one taxon in each of the four subtrees around it: edgesParts/makeTable, :866-889, :935-960)
// and the sum of their observed CF of the tree's resolution.  Such a quartet has a one-pair program (D[v], D[parent v]).
// bins[2*slot] = count, bins[2*slot+1] = sum in 2^-36 fixed point (64-bit integer atomics: order independent).
#define UBL_SCALE 68719476736.0
__global__ void __launch_bounds__(256) k_update_bl(SnaqTables T, const uint2 *__restrict__ prog2, const uint8_t *__restrict__ qhdr,
                                                   const uint32_t *__restrict__ perm, const double *__restrict__ obs, uint32_t nA1,
                                                   unsigned long long *__restrict__ bins) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nA1) return;
    const uint2 c = prog2[i];
    const unsigned nul = (unsigned)T.n_xe;
    if (c.y != nul * 0x10001u) return;
    const unsigned plus = c.x & 0xffffu, minus = c.x >> 16;
    const int first = T.n_xe + 1;
    if ((int)plus < first || (int)plus >= first + T.n_int || (int)minus < first || (int)minus >= first + T.n_int) return;
    const int v = T.dnode[plus - first], u = T.dnode[minus - first];
    if (T.parent[v] != u) return;                                  // a longer path that happens to need one pair
    const unsigned slot = T.pslot[v];
    if ((int)slot >= T.nparams) return;
    const unsigned h = qhdr[i];
    const double cf = obs[3 * (size_t)perm[i] + SNAQ_HDR_PERM(h)];
    atomicAdd(&bins[2 * slot], 1ull);
    atomicAdd(&bins[2 * slot + 1], (unsigned long long)__double2ll_rn(cf * UBL_SCALE));
}

// out[p] = -(sum over blocks of partial[b][p]), fixed order (logPseudoLik(d): src/pseudolik.jl:1417-1423).
// One block per candidate: 256 threads stride over the per-block partial sums with 4 independent
// accumulators each, then a fixed-order shuffle / shared-memory tree.
__global__ void __launch_bounds__(256) k_reduce_partials(const double *__restrict__ partial, int nblocks, int stride, int P,
                                                         double *__restrict__ out) {
    const int p = blockIdx.x, tid = threadIdx.x;
    if (p >= P) return;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    int b = tid;
    for (; b + 768 < nblocks; b += 1024) {
        s0 += partial[(size_t)b * stride + p];
        s1 += partial[(size_t)(b + 256) * stride + p];
        s2 += partial[(size_t)(b + 512) * stride + p];
        s3 += partial[(size_t)(b + 768) * stride + p];
    }
    for (; b < nblocks; b += 256) s0 += partial[(size_t)b * stride + p];
    double s = (s0 + s1) + (s2 + s3);
    __shared__ double red[8];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if ((tid & 31) == 0) red[tid >> 5] = s;
    __syncthreads();
    if (tid == 0) out[p] = -(((red[0] + red[1]) + (red[2] + red[3])) + ((red[4] + red[5]) + (red[6] + red[7])));
}
// first level of a two-level reduction over many rows (k_eval_runs writes one row per segment): block (x, y) adds rows
// [y * rows_per, (y + 1) * rows_per) for candidates [32 x, 32 x + 32) in a fixed order into out[y][.]
__global__ void __launch_bounds__(256) k_reduce_rows(const double *__restrict__ partial, int nrows, int rows_per, int stride, int P,
                                                     double *__restrict__ out) {
    __shared__ double sh[8][32];
    const int c = threadIdx.x & 31, r = threadIdx.x >> 5;
    const int p = blockIdx.x * 32 + c;
    const int r0 = blockIdx.y * rows_per, r1 = min(r0 + rows_per, nrows);
    double s = 0.0;
    if (p < P)
        for (int b = r0 + r; b < r1; b += 8) s += partial[(size_t)b * stride + p];
    sh[r][c] = s;
    __syncthreads();
    if (r == 0 && p < P) {
        double t = 0.0;
#pragma unroll
        for (int k = 0; k < 8; k++) t += sh[k][c];
        out[(size_t)blockIdx.y * stride + p] = t;
    }
}
// variant for the lanes kernel (candidates contiguous in a row of partial): a block of 32 x 8 threads handles 32
// candidates; thread (c, r) adds rows r, r+8, ... (coalesced 256-byte reads), then the 8 row groups are added
// in a fixed order
__global__ void __launch_bounds__(256) k_reduce_partials_t(const double *__restrict__ partial, int nblocks, int stride, int P,
                                                           double *__restrict__ out, const double *__restrict__ cst = nullptr) {
    __shared__ double sh[8][32];
    const int c = threadIdx.x & 31, r = threadIdx.x >> 5;
    const int p = blockIdx.x * 32 + c;
    double s = 0.0;
    grid_dependency_wait();                                      // the rows are the evaluation kernel's output
    if (p < P)
        for (int b = r; b < nblocks; b += 8) s += partial[(size_t)b * stride + p];
    sh[r][c] = s;
    __syncthreads();
    if (r == 0 && p < P) {
        double t = 0.0;
#pragma unroll
        for (int k = 0; k < 8; k++) t += sh[k][c];
        out[p] = cst ? -t - __ldg(cst) : -t;                 // cst = -(sum of the constants W3), see k_weights
    }
}

// --------------------------------------------------------------------------------------------
// FP64 roofline denominator: a dependent-chain DFMA microbenchmark (8 chains per thread)
// --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_dfma_peak(double *out, int iters, double a, double b) {
    double c0 = threadIdx.x, c1 = c0 + 1, c2 = c0 + 2, c3 = c0 + 3, c4 = c0 + 4, c5 = c0 + 5, c6 = c0 + 6, c7 = c0 + 7;
    for (int i = 0; i < iters; i++) {
        c0 = fma(c0, a, b); c1 = fma(c1, a, b); c2 = fma(c2, a, b); c3 = fma(c3, a, b);
        c4 = fma(c4, a, b); c5 = fma(c5, a, b); c6 = fma(c6, a, b); c7 = fma(c7, a, b);
    }
    double s = c0 + c1 + c2 + c3 + c4 + c5 + c6 + c7;
    if (s == 12345.6789) out[0] = s;   // never true in practice; keeps the chains alive
}

// order-dependent checksums of the device table (snaq_table_checksum): tests compare a patched table with a rebuilt one
__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__global__ void __launch_bounds__(256) k_checksum(const uint32_t *__restrict__ perm, const uint8_t *__restrict__ qhdr,
                                                  const uint32_t *__restrict__ off, const unsigned long long *__restrict__ prog8,
                                                  const unsigned long long *__restrict__ W8, int64_t nq, int64_t nchunks,
                                                  unsigned long long *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long a = 0, b = 0, c = 0;
    if (i < nq) {
        a = mix64((unsigned long long)i * 0x9E3779B97F4A7C15ull + (((unsigned long long)perm[i] << 32) | ((unsigned long long)qhdr[i] << 24)) + off[i]);
        for (int k = 0; k < 4; k++) c += mix64((unsigned long long)(4 * i + k) * 0x9E3779B97F4A7C15ull ^ W8[4 * i + k]);
    }
    for (int64_t j = i; j < nchunks; j += (int64_t)gridDim.x * blockDim.x) b += mix64((unsigned long long)j * 0x9E3779B97F4A7C15ull ^ prog8[j]);
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_down_sync(0xffffffffu, a, o); b += __shfl_down_sync(0xffffffffu, b, o); c += __shfl_down_sync(0xffffffffu, c, o);
    }
    if ((threadIdx.x & 31) == 0) { atomicAdd(out, a); atomicAdd(out + 1, b); atomicAdd(out + 2, c); }
}

// --------------------------------------------------------------------------------------------
// context
// --------------------------------------------------------------------------------------------
struct snaq_ctx {
    int device = 0, rank = 0, world = 1;
    void *comm = nullptr;                 // ncclComm_t of a quartet-sharded context (snaq_comm_init), else nullptr
    PeerMail *d_mail = nullptr;           // this rank's mailbox array [2][world] (fused epilogue of k_eval_direct)
    PeerMail *peer_mail[SNAQ_MAX_WORLD] = {nullptr};   // every rank's mailbox array as mapped into this process
    bool peer_ok = false;                 // the mailboxes of all peers are mapped (CUDA IPC over NVLink / PCIe peer access)
    unsigned long long comm_seq = 0;      // sequence number of the fused launches (the same on every rank)
    bool last_fused = false;              // the last k_eval_direct launch delivered totals, not partial sums
    double *d_red = nullptr; size_t red_cap = 0;       // device staging of values that are all-reduced before they go to the host
    cudaStream_t stream = nullptr;
    XArg xarg;                            // kernel-argument copy of a small host batch
    bool xarg_on = false;
    const double *xarg_src = nullptr;     // host batch being evaluated through kernel arguments (else nullptr)
    const double *xarg_base = nullptr;    // the (never dereferenced) device alias eval_device indexes for that batch
    uint32_t *status_out = nullptr;       // page-locked word the last block mirrors the status into (else nullptr)
    cudaStream_t copy_stream = nullptr;   // snaq_eval: host->device copies of the next chunk while the current one is evaluated
    cudaEvent_t copy_ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    std::string err;
    // data
    int ntaxa = 0;
    int64_t nq = 0, nq_total = 0, q_begin = 0;
    uint16_t *d_taxa = nullptr;
    double *d_obs = nullptr;
    uint8_t *d_sampled = nullptr;
    bool has_sampled = false;
    double *d_ci_lo = nullptr, *d_ci_hi = nullptr;   // snaq_set_ci: [nq][3] each
    bool has_ci = false;
    // network
    bool has_net = false;
    SnaqHostTables H;
    unsigned char *d_blob = nullptr;
    uint16_t *d_lca = nullptr;            // pair table of blob-tree LCAs (k_leaf_lca), ntaxa <= SNAQ_LCA_MAX_TAXA
    size_t lca_cap = 0;
    size_t blob_cap = 0;
    SnaqTables T{};
    double *d_xconst = nullptr;
    size_t xconst_cap = 0;
    uint32_t *d_off = nullptr;            // [nq+1] program offsets in 4-word chunks, class-sorted order
    uint16_t *d_prog = nullptr;
    size_t prog_cap = 0;
    uint64_t prog_words = 0;
    uint32_t maxlen = 0;                  // longest program, in chunks
    double *d_W = nullptr;
    double2 *d_WA = nullptr;              // [nq] (only [0,nA) used): class-A weight pairs streamed by k_eval_direct
    double *d_wpart = nullptr;            // per-block partial sums of the class-A constants W3
    double *d_negw3 = nullptr;            // -(sum of W3 over class A)
    uint32_t *d_perm = nullptr;           // [nq] sorted position -> original quartet
    uint8_t *d_qhdr = nullptr;            // [nq] header bytes, sorted order
    uint32_t *d_len4 = nullptr, *d_iota = nullptr;   // [nq] scratch of the build
    uint16_t *d_stage = nullptr;          // [nq][SNAQ_STAGE_WORDS] first words of every program, original order (k_build_count -> k_build_emit)
    uint8_t *d_hdr0 = nullptr;            // [nq] header bytes, original order
    uint8_t *d_dirty = nullptr; size_t dirty_cap = 0;   // snaq_patch_network: one byte per taxon
    int64_t patch_touched = -1;           // taxa re-described by the last snaq_patch_network (-1: everything)
    uint32_t *d_cls = nullptr, *d_cls_sorted = nullptr;   // sort keys (class, term hash)
    int64_t nA1 = 0, nA2 = 0, nA = 0, nB = 0, nC = 0;   // nA1: leading class-A quartets whose program is exactly one chunk (off[i] == i)
    // program deduplication (SNAQ_OPT_DEDUP): the compressed table the evaluation kernels read instead of the full one
    bool dedup = false;
    int64_t nU = 0, uA1 = 0, uA2 = 0, uA = 0, uB = 0, uC = 0;
    uint32_t *d_flags = nullptr, *d_uidx = nullptr;              // [nq]
    uint32_t *d_rstart = nullptr, *d_ulen = nullptr, *d_off_u = nullptr; size_t u_cap = 0;
    uint8_t *d_qhdr_u = nullptr;
    uint16_t *d_prog_u = nullptr; size_t prog_u_cap = 0; uint64_t prog_u_words = 0;
    double *d_W_u = nullptr; double2 *d_WA_u = nullptr; double *d_wpart_u = nullptr, *d_negw3_u = nullptr;
    double *d_pen_u = nullptr;            // [nU][4] per-run data of the -1e15 penalty (snaq_loglik_t5)
    // run table of the plain table (k_eval_runs): off_u / prog_u above hold one program per run, tdesc one descriptor per tile
    int runs_on = 1;                      // SNAQ_RUNS=0: the batched path uses the per-quartet kernels (k_eval_lanes / _p);
                                          // 1: run-length kernel when the runs are long enough to pay (RN_MIN_AVG_RUN); 2: always
    RunTile *d_tdesc = nullptr; size_t tdesc_cap = 0;
    uint32_t *d_rec_off = nullptr; size_t rec_off_cap = 0;       // [ntiles + 1] record offsets of the packed tile stream (16-byte units)
    unsigned char *d_pack = nullptr; size_t pack_cap = 0;         // the packed tile stream
    double *d_partial2 = nullptr; size_t partial2_cap = 0;        // second level of the segment reduction
    uint32_t *d_pidx = nullptr; size_t pidx_cap = 0;              // [ntiles + 1] first (tile, run) partial of every tile (k_wu_partials)
    double *d_wupart = nullptr; size_t wupart_cap = 0;            // [(tile, run) pairs][8] per-tile sums of the run weights
    bool sv_dedup = true;                 // single-vector paths (one objective call, gradient pass, optBL!) walk the compressed table
                                          // with run-summed weights (SNAQ_SV_DEDUP=0: the per-quartet table, as the batched path)
    uint32_t rn_ntiles = 0, rn_nseg = 0, rn_tps = 1, rn_ntAB = 0;
    bool pack_on = false;                 // the packed tile stream of k_eval_runs exists (per-quartet mode, SNAQ_RUNS != 0)
    uint32_t *d_status = nullptr;          // [0] build, [1] eval
    unsigned long long *d_trace = nullptr; // SNAQ_TRACE=1: [blocks][8] phase timestamps of the last k_eval_direct launch
    int trace_blocks = 0;
    unsigned int *d_ticket = nullptr;      // last-block ticket of k_eval_direct (reset by the kernel)
    unsigned long long *d_total = nullptr; // build statistics: total chunks, max chunks, class counts
    void *d_scan_tmp = nullptr;
    size_t scan_tmp_cap = 0;
    // scratch
    double *d_X = nullptr; size_t X_cap = 0;
    double *d_XF = nullptr; size_t XF_cap = 0;   // expanded parameter vectors [P][n_full]
    double *d_out = nullptr; size_t out_cap = 0;
    double *d_partial = nullptr; size_t partial_cap = 0;
    double *d_sq_w = nullptr, *d_sq_cum = nullptr, *d_sq_u = nullptr; long long *d_sq_idx = nullptr;   // snaq_sample_quartets
    size_t sq_w_cap = 0, sq_cum_cap = 0, sq_u_cap = 0, sq_idx_cap = 0;
    double *d_rb_e = nullptr, *d_rb_l = nullptr, *d_rb_d = nullptr; size_t rb_e_cap = 0, rb_l_cap = 0, rb_d_cap = 0;   // snaq_eval_quartets
    double *d_gpart = nullptr; size_t gpart_cap = 0;     // k_eval_grad_small: [blocks][rows] per-block gradient rows
    unsigned long long *d_gbins = nullptr; size_t gbins_cap = 0;   // k_eval_grad: [2][rows] grid-wide fixed-point gradient bins (zero between launches)
    double *d_gout = nullptr; size_t gout_cap = 0;       // [1 + n_full]: f, then d f / d xf
    double *h_pin = nullptr; size_t pin_cap = 0;
    int sm_count = 148;
    int eval_chunks = 2;                  // snaq_eval pipelines large batches in this many chunks (SNAQ_EVAL_CHUNKS = 1..4)
    int lanes_persistent = 1;             // SNAQ_LANES_PERSISTENT: 0 never, 1 for big xf tiles (default), 2 always
    bool reuse = true;                    // SNAQ_REUSE=0: the batched kernels evaluate every quartet in full (measurement aid)
    int lanes_chunk = 256;                // quartets per block of k_eval_lanes (SNAQ_LANES_CHUNK)
    int lanes_warps = 8;                  // warps per block of k_eval_lanes (SNAQ_LANES_WARPS = 8, 12 or 16)
    bool lanes_warps_set = false;
    bool pdl = true;                      // SNAQ_PDL=0: ordinary launches for the expand -> evaluate -> reduce chain
    int64_t counters[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    // snaq_optbl_topologies: worker contexts on the same device, each with a copy of this context's data
    uint64_t data_version = 1;            // bumped by snaq_set_data / snaq_set_obscf / snaq_set_sampled
    std::vector<snaq_ctx *> workers;
    std::vector<uint64_t> worker_version;
};

static thread_local std::string g_create_err;

struct nvtx_range {           // NVTX ranges for Nsight Systems / Compute timelines
    explicit nvtx_range(const char *name) { nvtxRangePushA(name); }
    ~nvtx_range() { nvtxRangePop(); }
};

#define CK(call)                                                                                          \
    do {                                                                                                  \
        cudaError_t e_ = (call);                                                                          \
        if (e_ != cudaSuccess) {                                                                          \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_) + " (snaq_engine.cu:" +         \
                       std::to_string(__LINE__) + ")";                                                    \
            return SNAQ_ECUDA;                                                                            \
        }                                                                                                 \
    } while (0)

// cudaFuncAttributeMaxDynamicSharedMemorySize is a property of the kernel, shared by every context and host thread of
// the process: raise it monotonically under a lock, so a launch prepared by one thread is never invalidated by a smaller
// value set for another context between its cudaFuncSetAttribute and its launch.
template <class K> static cudaError_t raise_smem_limit(K kernel, size_t smem) {
    struct Limit { const void *kernel; int device; size_t smem; };
    static std::mutex mu;
    static std::vector<Limit> limits;                        // per (kernel, device): the attribute lives in the device's module
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lock(mu);
    Limit *l = nullptr;
    for (auto &x : limits)
        if (x.kernel == (const void *)kernel && x.device == dev) l = &x;
    if (l && smem <= l->smem) return cudaSuccess;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (l) l->smem = smem;
    else limits.push_back({(const void *)kernel, dev, smem});
    return cudaSuccess;
}

// The single-call kernels finish by writing their results and then a status word with the DONE bit into page-locked
// host memory: the host spins on that word instead of waiting for the stream's completion semaphore (a few microseconds
// per objective call).  Every 4096 spins the stream is queried, so a failed launch or kernel surfaces as an error.
static cudaError_t wait_done(cudaStream_t stream, volatile uint32_t *flag) {
    for (unsigned spin = 1;; spin++) {
        if (*flag & SNAQ_S_DONE) break;
        if ((spin & 0xfffu) == 0) {
            cudaError_t e = cudaStreamQuery(stream);
            if (e == cudaSuccess) {                          // the kernel has ended: the flag is there, or something is wrong
                if (*flag & SNAQ_S_DONE) break;
                return cudaStreamSynchronize(stream) == cudaSuccess ? cudaErrorUnknown : cudaGetLastError();
            }
            if (e != cudaErrorNotReady) return e;
        }
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
        if (spin > 8192u && (spin & 0x3ffu) == 0) std::this_thread::yield();   // long kernel: let other host threads (workers) run
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    *flag &= ~SNAQ_S_DONE;
    return cudaSuccess;
}

template <class T> static int ensure(snaq_ctx *ctx, T *&p, size_t &cap, size_t need) {
    if (need <= cap && p) return SNAQ_OK;
    const bool regrow = p != nullptr;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    // a buffer that has to grow gets 25 % headroom: candidate topologies of a search differ a little in every size, and a
    // cudaFree + cudaMalloc pair costs more than the whole table build
    size_t n = std::max<size_t>(regrow ? need + need / 4 : need, 16);
    cudaError_t e = cudaMalloc((void **)&p, n * sizeof(T));
    if (e != cudaSuccess && n > need) { n = std::max<size_t>(need, 16); e = cudaMalloc((void **)&p, n * sizeof(T)); }
    if (e != cudaSuccess) { ctx->err = std::string("cudaMalloc: ") + cudaGetErrorString(e); return SNAQ_ENOMEM; }
    cap = n;
    return SNAQ_OK;
}
static int ensure_pin(snaq_ctx *ctx, size_t ndoubles) {
    if (ndoubles <= ctx->pin_cap && ctx->h_pin) return SNAQ_OK;
    if (ctx->h_pin) cudaFreeHost(ctx->h_pin);
    ctx->h_pin = nullptr;
    ctx->pin_cap = 0;
    size_t n = std::max<size_t>(ndoubles, 1024);
    cudaError_t e = cudaMallocHost((void **)&ctx->h_pin, n * sizeof(double));
    if (e != cudaSuccess) { ctx->err = std::string("cudaMallocHost: ") + cudaGetErrorString(e); return SNAQ_ENOMEM; }
    ctx->pin_cap = n;
    return SNAQ_OK;
}

static const char *build_err_text(uint32_t e) {
    switch (e) {
        case SNAQ_B_POLYTOMY: return "four taxa meet at a tree node: the network is not binary";
        case SNAQ_B_BLOCKS: return "block pattern of a quartet is inconsistent with the blob tree (taxon missing from the network, or not level-1)";
        case SNAQ_B_OVERFLOW: return "a quartet program exceeds the descriptor limits (path too long / too many cycles on one path)";
        default: return "unknown descriptor build error";
    }
}

static int status_to_error(snaq_ctx *ctx, uint32_t st) {
    if (!st) return SNAQ_OK;
    if (st & SNAQ_S_RANGE) { ctx->err = "new gamma / gammaz value should be between 0,1 and lengths nonnegative (src/snaq_optimization.jl:213,221,228; src/auxiliary.jl:119)"; return SNAQ_ERANGE; }
    if (st & SNAQ_S_NEGLEN) { ctx->err = "length can be negative, but not too negative (greater than -log(1.5)) or majorCF<0 (src/auxiliary.jl:120)"; return SNAQ_ENUMERIC; }
    if (st & SNAQ_S_COMM) { ctx->err = "quartet-sharded evaluation: a peer's partial sums did not arrive within 4 s (did every rank make the same call?)"; return SNAQ_ECUDA; }
    if (st & SNAQ_S_ZEROSUM) { ctx->err = "expCF not updated for quartet: sum(expCF)==0 (src/pseudolik.jl:1391)"; return SNAQ_ENUMERIC; }
    ctx->err = "numerical error in the evaluation kernel";
    return SNAQ_ENUMERIC;
}

// ---- NCCL, resolved at run time: the library has no link-time dependency on it (a single-GPU caller never needs it), and
// a process that already carries a libnccl.so.2 (e.g. PyTorch's) shares that copy -------------------------------------------
struct NcclApi {
    void *h = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
};
static NcclApi *nccl_api(std::string &err) {
    static std::mutex mu;
    static NcclApi api;
    static bool tried = false;
    std::lock_guard<std::mutex> lock(mu);
    if (!tried) {
        tried = true;
        const char *names[] = {getenv("SNAQ_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char *nm : names) {
            if (!nm || api.h) continue;
            api.h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        }
        if (api.h) {
            api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.h, "ncclGetUniqueId");
            api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.h, "ncclCommInitRank");
            api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.h, "ncclCommDestroy");
            api.AllReduce = (decltype(api.AllReduce))dlsym(api.h, "ncclAllReduce");
            api.AllGather = (decltype(api.AllGather))dlsym(api.h, "ncclAllGather");
            api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.h, "ncclGetErrorString");
        }
    }
    if (!api.h || !api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllReduce || !api.AllGather || !api.GetErrorString) {
        err = "NCCL (libnccl.so.2) could not be loaded: quartet-sharded contexts need it for their all-reduce (set SNAQ_NCCL_LIB to its path)";
        return nullptr;
    }
    return &api;
}
#define CKN(call)                                                                                              \
    do {                                                                                                       \
        ncclResult_t r_ = (call);                                                                              \
        if (r_ != ncclSuccess) {                                                                               \
            ctx->err = std::string(#call) + ": " + api->GetErrorString(r_);                                    \
            return SNAQ_ECUDA;                                                                                 \
        }                                                                                                      \
    } while (0)

// sum over the ranks of n doubles in device memory, in place, on `stream` (every rank of the communicator calls this)
static int comm_allreduce(snaq_ctx *ctx, double *d, size_t n, cudaStream_t stream) {
    NcclApi *api = nccl_api(ctx->err);
    if (!api) return SNAQ_ECUDA;
    CKN(api->AllReduce(d, d, n, ncclDouble, ncclSum, (ncclComm_t)ctx->comm, stream));
    return SNAQ_OK;
}

// the table an evaluation kernel walks: the full one, or the deduplicated one (one entry per distinct program)
struct EvalTable {
    const uint32_t *off; const uint16_t *prog; const uint8_t *qhdr; const double *W; const double2 *WA; const double *negw3;
    int64_t nq, nA1, nA2, nA, nB, nC;
    const double *pen;      // deduplicated table: per-run penalty data, else nullptr
};
static bool use_compressed(const snaq_ctx *ctx, bool full, bool single_vector) {
    return !full && ctx->nU > 0 && (ctx->dedup || (single_vector && ctx->sv_dedup));
}
static EvalTable eval_table(const snaq_ctx *ctx, bool full, bool single_vector = false) {
    if (use_compressed(ctx, full, single_vector)) return EvalTable{ctx->d_off_u, ctx->d_prog_u, ctx->d_qhdr_u, ctx->d_W_u, ctx->d_WA_u, ctx->d_negw3_u,
                                              ctx->nU, ctx->uA1, ctx->uA2, ctx->uA, ctx->uB, ctx->uC, ctx->d_pen_u};
    return EvalTable{ctx->d_off, ctx->d_prog, ctx->d_qhdr, ctx->d_W, ctx->d_WA, ctx->d_negw3, ctx->nq, ctx->nA1, ctx->nA2, ctx->nA, ctx->nB, ctx->nC,
                     nullptr};
}

static int run_weights(snaq_ctx *ctx) {
    if (!ctx->has_net || ctx->nq == 0) return SNAQ_OK;
    int64_t nb = (ctx->nq + 255) / 256;
    k_weights<<<(unsigned)nb, 256, 0, ctx->stream>>>(ctx->d_perm, ctx->d_qhdr, ctx->d_obs, ctx->has_sampled ? ctx->d_sampled : nullptr, ctx->nq,
                                                    ctx->nA1 + ctx->nA2, ctx->nA + ctx->nB, ctx->d_W, ctx->d_WA, ctx->d_wpart,
                                                    ctx->d_rec_off, ctx->rn_ntAB, (ctx->pack_on && ctx->rn_ntiles > 0) ? ctx->d_pack : nullptr);
    CK(cudaGetLastError());
    k_reduce_partials<<<2, 256, 0, ctx->stream>>>(ctx->d_wpart, (int)nb, 2, 2, ctx->d_negw3);
    CK(cudaGetLastError());
    ctx->counters[5] += 2;
    if (ctx->nU > 0 && ctx->rn_ntiles > 0) {
        // the run-summed weights of the compressed table (single-vector paths; SNAQ_OPT_DEDUP: every path).  No separate
        // class-A constant there (negw3_u stays 0): see launch_direct
        const uint32_t nt = ctx->rn_ntiles;
        k_wu_partials<<<(nt * 32 + 127) / 128, 128, 0, ctx->stream>>>(ctx->d_tdesc, ctx->d_pidx, reinterpret_cast<const double4 *>(ctx->d_W), nt,
                                                                      ctx->d_wupart);
        CK(cudaGetLastError());
        k_wu_combine<<<(unsigned)((ctx->nU * 32 + 127) / 128), 128, 0, ctx->stream>>>(ctx->d_rstart, ctx->d_tdesc, ctx->d_pidx, ctx->d_wupart, ctx->nU,
                                                                                     ctx->nq, ctx->nA + ctx->nB, ctx->rn_ntAB,
                                                                                     reinterpret_cast<double4 *>(ctx->d_W_u),
                                                                                     reinterpret_cast<double4 *>(ctx->d_pen_u));
        CK(cudaGetLastError());
        ctx->counters[5] += 2;
    }
    return SNAQ_OK;
}

// build the deduplicated table from the sorted full table (set_network, after k_build_emit)
static int build_dedup(snaq_ctx *ctx, int64_t pad_at) {
    const int64_t nq = ctx->nq;
    const uint2 *prog2 = reinterpret_cast<const uint2 *>(ctx->d_prog);
    k_dedup_flags<<<(unsigned)((nq + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_off, prog2, ctx->d_qhdr, nq, pad_at, ctx->d_flags);
    CK(cudaGetLastError());
    size_t t1 = ctx->scan_tmp_cap;
    CK(cub::DeviceScan::InclusiveSum(ctx->d_scan_tmp, t1, ctx->d_flags, ctx->d_uidx, nq, ctx->stream));
    // runs per class region: the run numbers at the region boundaries
    const int64_t bnd[5] = {ctx->nA1, ctx->nA1 + ctx->nA2, ctx->nA, ctx->nA + ctx->nB, nq};
    uint32_t ub[5] = {0, 0, 0, 0, 0};
    for (int k = 0; k < 5; k++)
        if (bnd[k] > 0) CK(cudaMemcpyAsync(&ub[k], ctx->d_uidx + (bnd[k] - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->uA1 = ub[0]; ctx->uA2 = (int64_t)ub[1] - ub[0]; ctx->uA = ub[2]; ctx->uB = (int64_t)ub[3] - ub[2]; ctx->uC = (int64_t)ub[4] - ub[3];
    ctx->nU = ub[4];
    const int64_t nu = ctx->nU;
    if ((size_t)nu + 1 > ctx->u_cap) {
        cudaFree(ctx->d_rstart); cudaFree(ctx->d_ulen); cudaFree(ctx->d_off_u); cudaFree(ctx->d_qhdr_u); cudaFree(ctx->d_W_u); cudaFree(ctx->d_WA_u);
        cudaFree(ctx->d_wpart_u); cudaFree(ctx->d_pen_u);
        ctx->d_rstart = ctx->d_ulen = ctx->d_off_u = nullptr; ctx->d_qhdr_u = nullptr; ctx->d_W_u = nullptr; ctx->d_WA_u = nullptr;
        ctx->d_wpart_u = nullptr; ctx->d_pen_u = nullptr;         // a failed cudaMalloc below must not leave dangling pointers
        ctx->u_cap = 0;
        const size_t n = (size_t)nu + 1 + (size_t)nu / 8;
        CK(cudaMalloc((void **)&ctx->d_rstart, (n + 1) * sizeof(uint32_t)));
        CK(cudaMalloc((void **)&ctx->d_ulen, (n + 1) * sizeof(uint32_t)));
        CK(cudaMalloc((void **)&ctx->d_off_u, (n + 8) * sizeof(uint32_t)));      // (+8: the tiles' aligned bulk copies read past the end)
        CK(cudaMalloc((void **)&ctx->d_qhdr_u, n + 1));
        CK(cudaMalloc((void **)&ctx->d_W_u, (n + 1) * 4 * sizeof(double)));
        CK(cudaMalloc((void **)&ctx->d_pen_u, (n + 1) * 4 * sizeof(double)));
        CK(cudaMalloc((void **)&ctx->d_WA_u, (n + 1) * sizeof(double2)));
        CK(cudaMalloc((void **)&ctx->d_wpart_u, ((n + 255) / 256 + 1) * sizeof(double)));
        ctx->u_cap = n;
    }
    if (!ctx->d_negw3_u) { CK(cudaMalloc((void **)&ctx->d_negw3_u, sizeof(double))); CK(cudaMemsetAsync(ctx->d_negw3_u, 0, sizeof(double), ctx->stream)); }
    const int64_t pad_u = (ctx->uA1 & 1) ? ctx->uA1 - 1 : -1;
    k_dedup_heads<<<(unsigned)((nq + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_flags, ctx->d_uidx, ctx->d_off, ctx->d_qhdr, nq, pad_at, pad_u,
                                                                        ctx->d_rstart, ctx->d_ulen, ctx->d_qhdr_u);
    CK(cudaGetLastError());
    CK(cudaMemsetAsync(ctx->d_off_u, 0, sizeof(uint32_t), ctx->stream));
    t1 = ctx->scan_tmp_cap;
    CK(cub::DeviceScan::InclusiveSum(ctx->d_scan_tmp, t1, ctx->d_ulen, ctx->d_off_u + 1, nu, ctx->stream));
    // no run is longer than the longest program (+ the one pad chunk): the bound saves a wait for the exact total
    ctx->prog_u_words = ((uint64_t)nu * ctx->maxlen + 1) * 4;
    int rc = ensure(ctx, ctx->d_prog_u, ctx->prog_u_cap, (size_t)ctx->prog_u_words + 8);
    if (rc) return rc;
    k_dedup_copy<<<(unsigned)((nu + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_rstart, ctx->d_off, prog2, ctx->d_off_u, nu, pad_u,
                                                                       (unsigned)ctx->H.n_xe, reinterpret_cast<uint2 *>(ctx->d_prog_u));
    CK(cudaGetLastError());
    ctx->counters[5] += 5;
    return SNAQ_OK;
}

// the packed tile stream of the run table (k_eval_runs); tiles never straddle the tree-like / 4-cycle boundary
static int build_run_tiles(snaq_ctx *ctx) {
    const int64_t nAB = ctx->nA + ctx->nB, nC = ctx->nq - nAB;
    const int64_t ntAB = (nAB + RN_TQ - 1) / RN_TQ, ntC = (nC + RN_TQC - 1) / RN_TQC;
    const int64_t nt = ntAB + ntC;
    if (nt >= 0x7fffffffLL) { ctx->err = "too many quartets on one GPU for the run table: shard them (world_size>1)"; return SNAQ_EINVAL; }
    ctx->rn_ntiles = (uint32_t)nt; ctx->rn_ntAB = (uint32_t)ntAB;
    ctx->rn_tps = (uint32_t)std::max<int64_t>(1, std::min<int64_t>(16, nt / 2048));       // depends on the table only
    ctx->rn_nseg = (uint32_t)((nt + ctx->rn_tps - 1) / ctx->rn_tps);
    int rc = ensure(ctx, ctx->d_tdesc, ctx->tdesc_cap, (size_t)nt + 1);
    if (rc) return rc;
    rc = ensure(ctx, ctx->d_rec_off, ctx->rec_off_cap, (size_t)nt + 2);
    if (rc) return rc;
    k_run_tiles<<<(unsigned)((nt + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_flags, ctx->d_uidx, ctx->d_off_u, ctx->nq, nAB, (uint32_t)ntAB,
                                                                      (uint32_t)nt, ctx->d_tdesc, ctx->d_rec_off + 1);
    CK(cudaGetLastError());
    size_t t1 = ctx->scan_tmp_cap;
    ctx->pack_on = ctx->runs_on && !ctx->dedup;
    if (ctx->pack_on) {
        // the packed stream k_eval_runs reads (per-quartet weights + programs per tile); the run table of the default mode
        // needs only the tile descriptors (run-summed weights)
        CK(cudaMemsetAsync(ctx->d_rec_off, 0, sizeof(uint32_t), ctx->stream));
        CK(cub::DeviceScan::InclusiveSum(ctx->d_scan_tmp, t1, ctx->d_rec_off + 1, ctx->d_rec_off + 1, nt, ctx->stream));
        uint32_t total16 = 0;
        CK(cudaMemcpyAsync(&total16, ctx->d_rec_off + nt, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        if ((uint64_t)nt * (RN_STAGE_BYTES / 16) >= 0xffffffffull) { ctx->err = "packed tile stream exceeds 64 GB: shard the quartets (world_size>1)"; return SNAQ_EINVAL; }
        rc = ensure(ctx, ctx->d_pack, ctx->pack_cap, (size_t)total16 * 16 + 16);
        if (rc) return rc;
        k_run_fill<<<(unsigned)((nt * 32 + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_tdesc, ctx->d_rec_off, ctx->d_off_u,
                                                                              reinterpret_cast<const uint2 *>(ctx->d_prog_u), (uint32_t)nt, ctx->d_pack);
        CK(cudaGetLastError());
    }
    // first (tile, run) partial of every tile: exclusive scan of the tiles' run counts; their total is at most nt + nU
    rc = ensure(ctx, ctx->d_pidx, ctx->pidx_cap, (size_t)nt + 2);
    if (rc) return rc;
    rc = ensure(ctx, ctx->d_wupart, ctx->wupart_cap, ((size_t)nt + (size_t)ctx->nU + 2) * 8);
    if (rc) return rc;
    k_tile_nruns<<<(unsigned)((nt + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_tdesc, (uint32_t)nt, ctx->d_pidx);
    CK(cudaGetLastError());
    t1 = ctx->scan_tmp_cap;
    CK(cub::DeviceScan::ExclusiveSum(ctx->d_scan_tmp, t1, ctx->d_pidx, ctx->d_pidx, nt + 1, ctx->stream));
    ctx->counters[5] += 5;
    return SNAQ_OK;
}

// the fused epilogue needs every rank to launch k_eval_direct (a rank without quartets launches nothing)
static bool fused_ok(const snaq_ctx *ctx) { return ctx->comm && ctx->peer_ok && ctx->nq_total >= ctx->world; }

template <int PC, bool OUT>
static int launch_direct(snaq_ctx *ctx, cudaStream_t stream, const double *dX, int Pc, double *dout, double *expcf, double *logpl,
                         double *deltacf) {
    const SnaqHostTables &H = ctx->H;
    if (ctx->xarg_src) {                       // host-resident candidates: this launch's slice travels as the kernel argument
        std::memcpy(ctx->xarg.v, ctx->xarg_src + (dX - ctx->xarg_base), sizeof(double) * (size_t)Pc * H.nparams);
        ctx->xarg_on = true;
    } else ctx->xarg_on = false;
    const size_t smem = 640 + 8 * PC * 8 + 8 * QD_SPAN * 8 + (OUT ? 0 : 8 * QD_NST * QD_STAGE_BYTES) +
                        (size_t)PC * (H.n_full + H.nparams) * 8;
    if (smem > 200 * 1024) { ctx->err = "network too large for the shared-memory tile of k_eval_direct"; return SNAQ_EINVAL; }
    const int per_sm = (int)std::min<size_t>(OUT ? 2 : 3, (225 * 1024) / (smem + 1024));
    EvalTable tb = eval_table(ctx, OUT, true);              // the per-quartet read-back walks the full table
    // compressed table: every entry through the generic path, which subtracts the run's own constant W3 entry by entry
    // (the separate class-A constant would turn the deviance into the difference of two large sums); the table is tiny
    if (use_compressed(ctx, OUT, true)) tb.nA1 = tb.nA2 = 0;
    const int64_t work = (tb.nq + 255) / 256;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(work, (int64_t)ctx->sm_count * per_sm));
    int rc = ensure(ctx, ctx->d_partial, ctx->partial_cap, (size_t)grid * PC);
    if (rc) return rc;
    if (ctx->d_trace) ctx->trace_blocks = grid;
    PeerBox box;
    std::memset(&box, 0, sizeof(box));
    ctx->last_fused = false;
    if (!OUT && fused_ok(ctx)) {                      // quartet-sharded: the all-reduce is fused into the kernel's epilogue
        for (int r = 0; r < ctx->world; r++) box.peer[r] = ctx->peer_mail[r];
        box.rank = ctx->rank; box.world = ctx->world;
        box.seq = ++ctx->comm_seq;
        ctx->last_fused = true;
    }
    CK(raise_smem_limit(k_eval_direct<PC, OUT>, smem));
    k_eval_direct<PC, OUT><<<grid, 256, smem, stream>>>(
        ctx->T, tb.off, reinterpret_cast<const uint2 *>(tb.prog), tb.qhdr, ctx->d_perm, tb.WA,
        reinterpret_cast<const double4 *>(tb.W), ctx->d_obs, (uint32_t)tb.nA1, (uint32_t)tb.nA2,
        (uint32_t)(tb.nA1 + (tb.nA1 & 1)), (uint32_t)tb.nq, dX,
        ctx->d_xconst, H.nh, H.nt, Pc, ctx->d_partial, dout, tb.negw3, ctx->d_ticket, expcf, logpl, deltacf, ctx->d_status + 1, ctx->d_trace,
        ctx->status_out, ctx->xarg_on ? 1 : 0, tb.pen, box, ctx->xarg);
    CK(cudaGetLastError());
    ctx->counters[0] += 1;
    ctx->counters[1] += ctx->nq * (int64_t)Pc;
    return SNAQ_OK;
}

// small batches (and the per-quartet read-back, one candidate at a time): at most 4 candidates per launch
static int eval_quartets_launch(snaq_ctx *ctx, int P, const double *dX, double *dout, cudaStream_t stream, double *expcf,
                                double *logpl, double *deltacf) {
    const SnaqHostTables &H = ctx->H;
    const bool per_quartet = expcf || logpl || deltacf;
    if (per_quartet) {
        // the per-quartet arrays describe the LAST candidate; the totals of the others come from the plain path
        if (P > 1) { int rc = eval_quartets_launch(ctx, P - 1, dX, dout, stream, nullptr, nullptr, nullptr); if (rc) return rc; }
        return launch_direct<1, true>(ctx, stream, dX + (size_t)(P - 1) * H.nparams, 1, dout + (P - 1), expcf, logpl, deltacf);
    }
    uint32_t *const status_out = ctx->status_out;
    for (int p0 = 0; p0 < P;) {
        const int left = P - p0;
        const double *dXp = dX + (size_t)p0 * H.nparams;
        int rc, Pc;
        ctx->status_out = left <= 4 ? status_out : nullptr;   // only the last launch raises the DONE flag the host polls
        if (left == 1) { Pc = 1; rc = launch_direct<1, false>(ctx, stream, dXp, 1, dout + p0, nullptr, nullptr, nullptr); }
        else if (left == 2) { Pc = 2; rc = launch_direct<2, false>(ctx, stream, dXp, 2, dout + p0, nullptr, nullptr, nullptr); }
        else { Pc = std::min(left, 4); rc = launch_direct<4, false>(ctx, stream, dXp, Pc, dout + p0, nullptr, nullptr, nullptr); }
        if (rc) { ctx->status_out = status_out; return rc; }
        p0 += Pc;
    }
    ctx->status_out = status_out;
    return SNAQ_OK;
}

// dispatch one evaluation of P candidates whose X is already on the device
static int eval_device_local(snaq_ctx *ctx, int P, const double *dX, double *dout, cudaStream_t stream,
                             double *expcf, double *logpl, double *deltacf) {
    const SnaqHostTables &H = ctx->H;
    if (ctx->nq == 0) { CK(cudaMemsetAsync(dout, 0, sizeof(double) * P, stream)); return SNAQ_OK; }
    const bool per_quartet = expcf || logpl || deltacf;
    const EvalTable tb = eval_table(ctx, per_quartet);
    const bool compressed_tb = use_compressed(ctx, per_quartet, false);
    const int64_t nq = tb.nq;
    const bool use_lanes = P >= LANES_MIN_P && !per_quartet;
    if (use_lanes) {
        if (ctx->pack_on && ctx->rn_ntiles > 0 &&
            (ctx->runs_on > 1 || ctx->nq >= (int64_t)RN_MIN_AVG_RUN * std::max<int64_t>(ctx->nU, 1))) {
            // run-length kernel: programs evaluated once per run of identical programs, weights streamed per quartet
            RunsMeta rm;
            rm.ntiles = ctx->rn_ntiles; rm.nseg = ctx->rn_nseg; rm.tps = ctx->rn_tps; rm.nul = (uint32_t)H.n_xe;
            rm.nA = ctx->nA; rm.noreuse = ctx->reuse ? 0 : 1;
            auto smem_of = [&](int warps, int cpl) {
                return (size_t)384 + ((((size_t)1 + warps * RN_NST) * 8 + 127) & ~size_t(127)) + (size_t)H.n_full * 256 * cpl +
                       (size_t)warps * RN_NST * RN_STAGE_BYTES;
            };
            // Two candidates per lane (the weights of a quartet are read from shared memory once for 64 candidates) with
            // 16-warp blocks when the wider xf tile fits; else one candidate per lane.  Measured at 100 taxa: 256 vectors
            // 0.40 ms either way (FP64 pipe 46 % against 36 % active), 4096 vectors 4.3 against 5.6 ms.
            // SNAQ_RUNS_CPL / SNAQ_RUNS_WARPS select the variants for measurements.
            int cpl = (P > 32 && smem_of(16, 2) <= 224 * 1024) ? 2 : 1;
            if (const char *ev = getenv("SNAQ_RUNS_CPL")) { const int v = atoi(ev); if (v == 1 || (v == 2 && P > 32 && smem_of(8, 2) <= 220 * 1024)) cpl = v; }
            int warps = (smem_of(16, cpl) <= 224 * 1024 && (smem_of(8, cpl) > 110 * 1024 || (int64_t)ctx->rn_nseg >= 16 * (int64_t)ctx->sm_count)) ? 16 : 8;
            if (const char *ev = getenv("SNAQ_RUNS_WARPS")) { const int v = atoi(ev); if (v == 8 || (v == 16 && smem_of(16, cpl) <= 224 * 1024)) warps = v; }
            const size_t smem = smem_of(warps, cpl);
            if (smem > 227 * 1024) { ctx->err = "network too large for the shared-memory tile of k_eval_runs"; return SNAQ_EINVAL; }
            const int tw = 32 * cpl;
            const int ngr = (P + tw - 1) / tw, Pp = ngr * tw;
            if (ngr > 65535) { ctx->err = "batch too large for one launch"; return SNAQ_EINVAL; }
            { int rcx = ensure(ctx, ctx->d_XF, ctx->XF_cap, (size_t)Pp * H.n_full); if (rcx) return rcx; }
            {
                const int64_t totx = (int64_t)Pp * H.n_full;
                k_expand_x<<<(unsigned)((totx + 255) / 256), 256, 0, stream>>>(ctx->T, dX, ctx->d_xconst, P, H.nh, H.nt, tw, ctx->d_XF,
                                                                                ctx->d_status + 1);
                CK(cudaGetLastError());
            }
            int per_sm = 1;
#define OCC_RUNS(WN, CP) do { CK(raise_smem_limit(k_eval_runs<WN, CP>, smem)); CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_eval_runs<WN, CP>, WN * 32, smem)); } while (0)
            if (warps == 16 && cpl == 2) OCC_RUNS(16, 2);
            else if (warps == 16) OCC_RUNS(16, 1);
            else if (cpl == 2) OCC_RUNS(8, 2);
            else OCC_RUNS(8, 1);
#undef OCC_RUNS
            per_sm = std::max(per_sm, 1);
            const int64_t slots = (int64_t)ctx->sm_count * per_sm;
            const int64_t want = ((int64_t)rm.nseg + warps - 1) / warps;              // blocks that still have a segment per warp
            const int gx = (int)std::max<int64_t>(1, std::min<int64_t>(want, slots / ngr));
            const int S = (int)std::max<int64_t>(1, std::min<int64_t>(64, rm.nseg / 32));    // depends on the table only
            const int rows_per = ((int)rm.nseg + S - 1) / S;
            int rc = ensure(ctx, ctx->d_partial, ctx->partial_cap, (size_t)rm.nseg * Pp);
            if (rc) return rc;
            rc = ensure(ctx, ctx->d_partial2, ctx->partial2_cap, (size_t)S * Pp);
            if (rc) return rc;
            dim3 grid((unsigned)gx, (unsigned)ngr);
            const uint2 *pu = reinterpret_cast<const uint2 *>(ctx->d_prog_u);
#define LAUNCH_RUNS(WN, CP)                                                                                                    \
    do {                                                                                                                       \
        CK(raise_smem_limit(k_eval_runs<WN, CP>, smem));                                                                         \
        if (ctx->pdl) CK(launch_pdl(k_eval_runs<WN, CP>, grid, dim3(WN * 32), smem, stream, (const unsigned char *)ctx->d_pack,    \
                                    (const uint32_t *)ctx->d_rec_off, pu, rm, (const double *)ctx->d_XF, H.n_full, ctx->d_partial, Pp, \
                                    ctx->d_status + 1));                                                                       \
        else k_eval_runs<WN, CP><<<grid, WN * 32, smem, stream>>>(ctx->d_pack, ctx->d_rec_off, pu, rm, ctx->d_XF, H.n_full,     \
                                                                 ctx->d_partial, Pp, ctx->d_status + 1);                        \
    } while (0)
            if (warps == 16 && cpl == 2) LAUNCH_RUNS(16, 2);
            else if (warps == 16) LAUNCH_RUNS(16, 1);
            else if (cpl == 2) LAUNCH_RUNS(8, 2);
            else LAUNCH_RUNS(8, 1);
#undef LAUNCH_RUNS
            CK(cudaGetLastError());
            if (S > 1) {
                k_reduce_rows<<<dim3((unsigned)((P + 31) / 32), (unsigned)S), 256, 0, stream>>>(ctx->d_partial, (int)rm.nseg, rows_per, Pp, P,
                                                                                                  ctx->d_partial2);
                CK(cudaGetLastError());
                k_reduce_partials_t<<<(P + 31) / 32, 256, 0, stream>>>(ctx->d_partial2, S, Pp, P, dout, ctx->d_negw3 + 1);
            } else {
                k_reduce_partials_t<<<(P + 31) / 32, 256, 0, stream>>>(ctx->d_partial, (int)rm.nseg, Pp, P, dout, ctx->d_negw3 + 1);
            }
            CK(cudaGetLastError());
            ctx->counters[0] += 1;
            ctx->counters[1] += ctx->nq * (int64_t)P;
            return SNAQ_OK;
        }
        const int Prows = (P + 31) & ~31;
        int rc0 = ensure(ctx, ctx->d_XF, ctx->XF_cap, (size_t)Prows * H.n_full);
        if (rc0) return rc0;
        const int64_t tot = (int64_t)Prows * H.n_full;
        k_expand_x<<<(unsigned)((tot + 255) / 256), 256, 0, stream>>>(ctx->T, dX, ctx->d_xconst, P, H.nh, H.nt, 32, ctx->d_XF,
                                                                      ctx->d_status + 1);
        CK(cudaGetLastError());
        const double *dXF = ctx->d_XF;
        const int ngroups = (P + 31) / 32, Ppad = ngroups * 32;
        ClassRanges cr;
        cr.nA1 = tb.nA1; cr.nA = tb.nA; cr.nB = tb.nB; cr.nC = tb.nC;
        cr.nA2 = tb.nA2; cr.a2base = (uint32_t)(tb.nA1 + (tb.nA1 & 1));
        cr.noreuse = ctx->reuse ? 0 : 1;
        cr.pen = tb.pen;
        if (ngroups > 65535) { ctx->err = "batch too large for one launch"; return SNAQ_EINVAL; }
        // xf tiles of 48 KB and more (~100 taxa and up): reloading the tile for every chunk of quartets is most of the
        // L2 traffic, and from ~100 KB only one block fits per SM: the persistent variant keeps the tile resident
        // (200 taxa, P = 64: 31.2 -> 16.9 ms; 100 taxa, P = 256: 339 -> 383 G evals/s).  With small tiles the
        // one-block-per-chunk kernel is faster (config 2: 329 vs 254 G evals/s): more, independent blocks per SM.
        const bool big_tile = lanesp_smem_layout(H.n_full, 64, 8, ctx->maxlen).total > 110 * 1024;
        // ~100 taxa: the tile reload is most of the L2 traffic -- of a long table; a run table of a few thousand programs is
        // faster with one block per (chunk, group) (100 taxa, 9,419 programs, P = 256: 0.043 vs 0.052 ms per step)
        const bool mid_tile = (size_t)H.n_full * 256 >= 48 * 1024 && nq >= 262144;
        if (ctx->lanes_persistent == 1 ? (big_tile || mid_tile) : ctx->lanes_persistent > 1) {
            // persistent blocks: the xf tile stays in shared memory, chunks stream through two TMA stages; 16 warps per block
            int warps = big_tile ? 16 : ctx->lanes_warps, chunk = 256;
            while (chunk > 32 && lanesp_smem_layout(H.n_full, chunk, warps, ctx->maxlen).total > 200 * 1024) chunk >>= 1;
            const size_t smem = lanesp_smem_layout(H.n_full, chunk, warps, ctx->maxlen).total;
            if (smem > 227 * 1024) { ctx->err = "network too large for the shared-memory tile of k_eval_lanes"; return SNAQ_EINVAL; }
            cr.chA = (int)((cr.nA + chunk - 1) / chunk); cr.chB = (int)((cr.nB + chunk - 1) / chunk); cr.chC = (int)((cr.nC + chunk - 1) / chunk);
            const int64_t nchunks = (int64_t)cr.chA + cr.chB + cr.chC;
            if (nchunks > 2147483647LL) { ctx->err = "batch too large for one launch"; return SNAQ_EINVAL; }
            // blocks per group: about two rounds of resident blocks over the whole grid
            const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(warps == 8 ? 4 : (warps == 12 ? 3 : 2), (225 * 1024) / (smem + 1024)));
            const int64_t slots = (int64_t)ctx->sm_count * per_sm;
            const int LANESP_SUPER = (int)std::max<int64_t>(1, std::min<int64_t>(8, nchunks / 1024));   // depends on the table only, not on P
            const int64_t nsuper = (nchunks + LANESP_SUPER - 1) / LANESP_SUPER;
            const int gx = (int)std::max<int64_t>(1, std::min<int64_t>(nsuper, (2 * slots) / ngroups));
            int rc = ensure(ctx, ctx->d_partial, ctx->partial_cap, (size_t)nsuper * Ppad);
            if (rc) return rc;
            dim3 grid((unsigned)gx, (unsigned)ngroups);
#define LAUNCH_LANESP(WN)                                                                                                      \
    do {                                                                                                                       \
        CK(raise_smem_limit(k_eval_lanes_p<WN>, smem));                                                                          \
        if (ctx->pdl) CK(launch_pdl(k_eval_lanes_p<WN>, grid, dim3(WN * 32), smem, stream, tb.off, tb.prog, tb.qhdr, tb.W, cr, chunk, \
                                    ctx->maxlen, dXF, H.n_full, ctx->d_partial, Ppad, ctx->d_status + 1, LANESP_SUPER));        \
        else k_eval_lanes_p<WN><<<grid, WN * 32, smem, stream>>>(tb.off, tb.prog, tb.qhdr, tb.W, cr, chunk, ctx->maxlen,        \
                                                                dXF, H.n_full, ctx->d_partial, Ppad, ctx->d_status + 1, LANESP_SUPER); \
    } while (0)
            if (warps == 8) LAUNCH_LANESP(8);
            else if (warps == 12) LAUNCH_LANESP(12);
            else LAUNCH_LANESP(16);
#undef LAUNCH_LANESP
            CK(cudaGetLastError());
            if (ctx->pdl) CK(launch_pdl(k_reduce_partials_t, dim3((P + 31) / 32), dim3(256), 0, stream, (const double *)ctx->d_partial, (int)nsuper,
                                        Ppad, P, dout, (const double *)nullptr));
            else k_reduce_partials_t<<<(P + 31) / 32, 256, 0, stream>>>(ctx->d_partial, (int)nsuper, Ppad, P, dout);
            CK(cudaGetLastError());
            ctx->counters[0] += 1;
            ctx->counters[1] += ctx->nq * (int64_t)P;
            return SNAQ_OK;
        }
        // one block per (chunk, group): quartets per block.  Enough blocks to fill the machine, bounded by shared memory.
        // run table (every entry a different program: no reuse between neighbours, long dependent chains of exp / log): 16 warps
        const int warps = (compressed_tb && !ctx->lanes_warps_set) ? 16 : ctx->lanes_warps;
        int chunk = ctx->lanes_chunk;
        while (chunk > 64 && (int64_t)ngroups * ((nq + chunk - 1) / chunk) < (int64_t)ctx->sm_count * 8) chunk >>= 1;
        while (chunk > 32 && lanes_smem_layout(H.n_full, chunk, warps, ctx->maxlen).total > 200 * 1024) chunk >>= 1;
        const size_t smem = lanes_smem_layout(H.n_full, chunk, warps, ctx->maxlen).total;
        if (smem > 227 * 1024) { ctx->err = "network too large for the shared-memory tile of k_eval_lanes"; return SNAQ_EINVAL; }
        cr.chA = (int)((cr.nA + chunk - 1) / chunk); cr.chB = (int)((cr.nB + chunk - 1) / chunk); cr.chC = (int)((cr.nC + chunk - 1) / chunk);
        const int64_t nchunks = (int64_t)cr.chA + cr.chB + cr.chC;
        if (nchunks > 2147483647LL) { ctx->err = "batch too large for one launch"; return SNAQ_EINVAL; }
        int rc = ensure(ctx, ctx->d_partial, ctx->partial_cap, (size_t)nchunks * Ppad);
        if (rc) return rc;
        dim3 grid((unsigned)nchunks, (unsigned)ngroups);
#define LAUNCH_LANES(WN)                                                                                                       \
    do {                                                                                                                       \
        CK(raise_smem_limit(k_eval_lanes<WN>, smem));                                                                            \
        if (ctx->pdl) CK(launch_pdl(k_eval_lanes<WN>, grid, dim3(WN * 32), smem, stream, tb.off, tb.prog, tb.qhdr, tb.W, cr, chunk, \
                                    ctx->maxlen, dXF, H.n_full, ctx->d_partial, Ppad, ctx->d_status + 1));                     \
        else k_eval_lanes<WN><<<grid, WN * 32, smem, stream>>>(tb.off, tb.prog, tb.qhdr, tb.W, cr, chunk, ctx->maxlen, dXF,    \
                                                              H.n_full, ctx->d_partial, Ppad, ctx->d_status + 1);              \
    } while (0)
        if (warps == 8) LAUNCH_LANES(8);
        else if (warps == 12) LAUNCH_LANES(12);
        else LAUNCH_LANES(16);
#undef LAUNCH_LANES
        CK(cudaGetLastError());
        if (ctx->pdl) CK(launch_pdl(k_reduce_partials_t, dim3((P + 31) / 32), dim3(256), 0, stream, (const double *)ctx->d_partial, (int)nchunks,
                                    Ppad, P, dout, (const double *)nullptr));
        else k_reduce_partials_t<<<(P + 31) / 32, 256, 0, stream>>>(ctx->d_partial, (int)nchunks, Ppad, P, dout);
        CK(cudaGetLastError());
        ctx->counters[0] += 1;
        ctx->counters[1] += ctx->nq * (int64_t)P;
        return SNAQ_OK;
    }
    // small batches: thread per quartet, up to 4 candidates per launch
    return eval_quartets_launch(ctx, P, dX, dout, stream, expcf, logpl, deltacf);
}

// One evaluation of P candidates; on a quartet-sharded context with a communicator the values are the totals over all
// ranks: small batches get them from the fused epilogue of k_eval_direct, everything else from an ncclAllReduce of the
// P partial sums on the same stream (dout must then be device memory).
static int eval_device(snaq_ctx *ctx, int P, const double *dX, double *dout, cudaStream_t stream,
                       double *expcf, double *logpl, double *deltacf) {
    ctx->last_fused = false;
    int rc = eval_device_local(ctx, P, dX, dout, stream, expcf, logpl, deltacf);
    if (rc) return rc;
    if (ctx->comm && !(expcf || logpl || deltacf) && !ctx->last_fused) return comm_allreduce(ctx, dout, (size_t)P, stream);
    return SNAQ_OK;
}

// --------------------------------------------------------------------------------------------
// C ABI
// --------------------------------------------------------------------------------------------
extern "C" {

const char *snaq_version(void) { return "snaq_b200 0.1 (sm_100a)"; }

const char *snaq_last_error(const snaq_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

int snaq_create(const snaq_opts *opts, snaq_ctx **out) {
    if (!out) return SNAQ_EINVAL;
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        g_create_err = std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                       " (this engine has no CPU path)";
        return SNAQ_ECUDA;
    }
    snaq_ctx *ctx = new snaq_ctx();
    ctx->device = opts ? opts->device : 0;
    ctx->rank = opts ? opts->rank : 0;
    ctx->world = (opts && opts->world_size > 0) ? opts->world_size : 1;
    // every path walks the run table with run-summed weights unless the caller asks for one FMA pair per (quartet, vector)
    ctx->dedup = !(opts && (opts->flags & SNAQ_OPT_PER_QUARTET));
    if (const char *ev = getenv("SNAQ_DEDUP")) ctx->dedup = atoi(ev) != 0;
    ctx->sv_dedup = ctx->dedup;            // single-vector paths follow (SNAQ_SV_DEDUP overrides, below)
    if (ctx->device < 0 || ctx->device >= ndev || ctx->rank < 0 || ctx->rank >= ctx->world) {
        g_create_err = "bad device ordinal or rank";
        delete ctx;
        return SNAQ_EINVAL;
    }
    auto fail = [&](cudaError_t ce, const char *what) {
        g_create_err = std::string(what) + ": " + cudaGetErrorString(ce);
        delete ctx;
        return SNAQ_ECUDA;
    };
    if ((e = cudaSetDevice(ctx->device)) != cudaSuccess) return fail(e, "cudaSetDevice");
    if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) return fail(e, "cudaStreamCreate");
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, ctx->device)) != cudaSuccess) return fail(e, "cudaGetDeviceProperties");
    ctx->sm_count = prop.multiProcessorCount;
    if (const char *ev = getenv("SNAQ_TRACE")) {
        if (atoi(ev) > 0 && cudaMalloc((void **)&ctx->d_trace, (size_t)ctx->sm_count * 4 * 8 * sizeof(unsigned long long)) != cudaSuccess) ctx->d_trace = nullptr;
    }
    if (const char *ev = getenv("SNAQ_EVAL_CHUNKS")) { int v = atoi(ev); if (v >= 1 && v <= 4) ctx->eval_chunks = v; }
    if (const char *ev = getenv("SNAQ_REUSE")) ctx->reuse = atoi(ev) != 0;
    if (const char *ev = getenv("SNAQ_RUNS")) ctx->runs_on = atoi(ev);
    if (const char *ev = getenv("SNAQ_SV_DEDUP")) ctx->sv_dedup = atoi(ev) != 0;
    if (const char *ev = getenv("SNAQ_LANES_CHUNK")) { int v = atoi(ev); if (v == 128 || v == 256 || v == 512 || v == 1024) ctx->lanes_chunk = v; }
    if (const char *ev = getenv("SNAQ_LANES_PERSISTENT")) ctx->lanes_persistent = atoi(ev);
    if (const char *ev = getenv("SNAQ_PDL")) ctx->pdl = atoi(ev) != 0;
    if (const char *ev = getenv("SNAQ_LANES_WARPS")) { int v = atoi(ev); if (v == 8 || v == 12 || v == 16) { ctx->lanes_warps = v; ctx->lanes_warps_set = true; } }
    if ((e = cudaMalloc((void **)&ctx->d_status, 4 * sizeof(uint32_t))) != cudaSuccess) return fail(e, "cudaMalloc");
    if ((e = cudaMalloc((void **)&ctx->d_total, 8 * sizeof(unsigned long long))) != cudaSuccess) return fail(e, "cudaMalloc");
    cudaMemset(ctx->d_status, 0, 4 * sizeof(uint32_t));
    if ((e = cudaMalloc((void **)&ctx->d_ticket, sizeof(unsigned int))) != cudaSuccess) return fail(e, "cudaMalloc");
    cudaMemset(ctx->d_ticket, 0, sizeof(unsigned int));
    *out = ctx;
    return SNAQ_OK;
}

int snaq_destroy(snaq_ctx *ctx) {
    if (!ctx) return SNAQ_OK;
    for (snaq_ctx *w : ctx->workers) snaq_destroy(w);
    ctx->workers.clear();
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->comm) {
        for (int rk = 0; rk < ctx->world; rk++)
            if (rk != ctx->rank && ctx->peer_mail[rk]) cudaIpcCloseMemHandle(ctx->peer_mail[rk]);
        std::string e;
        if (NcclApi *api = nccl_api(e)) api->CommDestroy((ncclComm_t)ctx->comm);
        ctx->comm = nullptr;
    }
    cudaFree(ctx->d_mail); cudaFree(ctx->d_red);
    cudaFree(ctx->d_ci_lo); cudaFree(ctx->d_ci_hi);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    cudaFree(ctx->d_taxa); cudaFree(ctx->d_obs); cudaFree(ctx->d_sampled); cudaFree(ctx->d_blob); cudaFree(ctx->d_lca); cudaFree(ctx->d_xconst);
    cudaFree(ctx->d_off); cudaFree(ctx->d_prog); cudaFree(ctx->d_W); cudaFree(ctx->d_WA); cudaFree(ctx->d_wpart); cudaFree(ctx->d_negw3); cudaFree(ctx->d_status); cudaFree(ctx->d_total); cudaFree(ctx->d_ticket); cudaFree(ctx->d_trace);
    cudaFree(ctx->d_flags); cudaFree(ctx->d_uidx); cudaFree(ctx->d_rstart); cudaFree(ctx->d_ulen); cudaFree(ctx->d_off_u); cudaFree(ctx->d_qhdr_u);
    cudaFree(ctx->d_stage); cudaFree(ctx->d_hdr0); cudaFree(ctx->d_dirty); cudaFree(ctx->d_pidx); cudaFree(ctx->d_wupart);
    cudaFree(ctx->d_tdesc); cudaFree(ctx->d_rec_off); cudaFree(ctx->d_pack); cudaFree(ctx->d_partial2);
    cudaFree(ctx->d_prog_u); cudaFree(ctx->d_W_u); cudaFree(ctx->d_WA_u); cudaFree(ctx->d_wpart_u); cudaFree(ctx->d_negw3_u); cudaFree(ctx->d_pen_u);
    cudaFree(ctx->d_scan_tmp); cudaFree(ctx->d_X); cudaFree(ctx->d_XF); cudaFree(ctx->d_out); cudaFree(ctx->d_partial); cudaFree(ctx->d_gbins); cudaFree(ctx->d_gpart); cudaFree(ctx->d_rb_e); cudaFree(ctx->d_rb_l); cudaFree(ctx->d_rb_d); cudaFree(ctx->d_sq_w); cudaFree(ctx->d_sq_cum); cudaFree(ctx->d_sq_u); cudaFree(ctx->d_sq_idx); cudaFree(ctx->d_gout);
    cudaFree(ctx->d_perm); cudaFree(ctx->d_qhdr); cudaFree(ctx->d_len4); cudaFree(ctx->d_iota); cudaFree(ctx->d_cls);
    cudaFree(ctx->d_cls_sorted);
    if (ctx->h_pin) cudaFreeHost(ctx->h_pin);
    if (ctx->copy_stream) { for (int k = 0; k < 8; k++) cudaEventDestroy(ctx->copy_ev[k]); cudaStreamDestroy(ctx->copy_stream); }
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return SNAQ_OK;
}


// ---- quartet-sharded mode: the collective lives inside the library ---------------------------------------------------
int snaq_comm_unique_id(void *id128) {
    if (!id128) return SNAQ_EINVAL;
    NcclApi *api = nccl_api(g_create_err);
    if (!api) return SNAQ_ECUDA;
    ncclUniqueId id;
    ncclResult_t r = api->GetUniqueId(&id);
    if (r != ncclSuccess) { g_create_err = std::string("ncclGetUniqueId: ") + api->GetErrorString(r); return SNAQ_ECUDA; }
    static_assert(sizeof(id) == 128, "ncclUniqueId");
    std::memcpy(id128, &id, sizeof(id));
    return SNAQ_OK;
}

int snaq_comm_init(snaq_ctx *ctx, const void *id128) {
    if (!ctx || !id128) return SNAQ_EINVAL;
    if (ctx->world <= 1) return SNAQ_OK;                                  // nothing to combine
    if (ctx->comm) { ctx->err = "snaq_comm_init: the context already has a communicator"; return SNAQ_EINVAL; }
    if (ctx->world > SNAQ_MAX_WORLD) { ctx->err = "snaq_comm_init: at most " + std::to_string(SNAQ_MAX_WORLD) + " ranks"; return SNAQ_EINVAL; }
    NcclApi *api = nccl_api(ctx->err);
    if (!api) return SNAQ_ECUDA;
    CK(cudaSetDevice(ctx->device));
    ncclUniqueId id;
    std::memcpy(&id, id128, sizeof(id));
    ncclComm_t comm = nullptr;
    CKN(api->CommInitRank(&comm, ctx->world, id, ctx->rank));
    ctx->comm = comm;
    // mailboxes for the fused epilogue of k_eval_direct: every rank's array [2][world] mapped into every rank (CUDA IPC).
    // If a mapping fails (no peer access between two devices) the small batches fall back to ncclAllReduce as well.
    const size_t nmail = 2 * (size_t)ctx->world;
    CK(cudaMalloc((void **)&ctx->d_mail, nmail * sizeof(PeerMail)));
    CK(cudaMemset(ctx->d_mail, 0, nmail * sizeof(PeerMail)));
    cudaIpcMemHandle_t mine;
    CK(cudaIpcGetMemHandle(&mine, ctx->d_mail));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t");
    unsigned char *d_h = nullptr;
    CK(cudaMalloc((void **)&d_h, 64 * ((size_t)ctx->world + 1)));
    CK(cudaMemcpy(d_h + 64 * (size_t)ctx->world, &mine, 64, cudaMemcpyHostToDevice));
    ncclResult_t r = api->AllGather(d_h + 64 * (size_t)ctx->world, d_h, 64, ncclChar, comm, ctx->stream);
    if (r != ncclSuccess) { cudaFree(d_h); ctx->err = std::string("ncclAllGather: ") + api->GetErrorString(r); return SNAQ_ECUDA; }
    std::vector<cudaIpcMemHandle_t> all((size_t)ctx->world);
    cudaError_t e = cudaMemcpyAsync(all.data(), d_h, 64 * (size_t)ctx->world, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_h);
    if (e != cudaSuccess) { ctx->err = std::string("snaq_comm_init: ") + cudaGetErrorString(e); return SNAQ_ECUDA; }
    int ok = getenv("SNAQ_NO_PEER_EPILOGUE") ? 0 : 1;
    for (int rk = 0; rk < ctx->world && ok; rk++) {
        if (rk == ctx->rank) { ctx->peer_mail[rk] = ctx->d_mail; continue; }
        void *pp = nullptr;
        if (cudaIpcOpenMemHandle(&pp, all[(size_t)rk], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; break; }
        ctx->peer_mail[rk] = (PeerMail *)pp;
    }
    // every rank must take the same path: the fused epilogue is used only if ALL ranks mapped ALL mailboxes
    int *d_ok = nullptr;
    CK(cudaMalloc((void **)&d_ok, sizeof(int)));
    CK(cudaMemcpy(d_ok, &ok, sizeof(int), cudaMemcpyHostToDevice));
    r = api->AllReduce(d_ok, d_ok, 1, ncclInt, ncclMin, comm, ctx->stream);
    if (r == ncclSuccess) { e = cudaMemcpyAsync(&ok, d_ok, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream); if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream); }
    cudaFree(d_ok);
    if (r != ncclSuccess) { ctx->err = std::string("ncclAllReduce: ") + api->GetErrorString(r); return SNAQ_ECUDA; }
    if (e != cudaSuccess) { ctx->err = std::string("snaq_comm_init: ") + cudaGetErrorString(e); return SNAQ_ECUDA; }
    ctx->peer_ok = ok != 0;
    ctx->comm_seq = 0;
    return SNAQ_OK;
}

int snaq_comm_info(snaq_ctx *ctx, int32_t *rank, int32_t *world_size, int32_t *has_comm, int32_t *fused_epilogue) {
    if (!ctx) return SNAQ_EINVAL;
    if (rank) *rank = ctx->rank;
    if (world_size) *world_size = ctx->world;
    if (has_comm) *has_comm = ctx->comm ? 1 : 0;
    if (fused_epilogue) *fused_epilogue = fused_ok(ctx) ? 1 : 0;
    return SNAQ_OK;
}

int snaq_set_data(snaq_ctx *ctx, int32_t ntaxa, int64_t nq, const uint16_t *taxa, const double *obscf) {
    if (!ctx) return SNAQ_EINVAL;
    if (ntaxa < 4 || ntaxa > 65535 || nq < 0 || (nq > 0 && (!taxa || !obscf))) { ctx->err = "bad arguments to snaq_set_data"; return SNAQ_EINVAL; }
    CK(cudaSetDevice(ctx->device));
    for (int64_t i = 0; i < nq * 4; i++)
        if (taxa[i] >= ntaxa) { ctx->err = "taxon id out of range in quartet " + std::to_string(i / 4); return SNAQ_EINVAL; }
    for (int64_t i = 0; i < nq * 3; i++)
        if (!(obscf[i] >= 0.0 && obscf[i] <= 1.0)) { ctx->err = "obsCF must be between (0,1) in quartet " + std::to_string(i / 3); return SNAQ_EINVAL; }
    // quartet-sharded mode: this rank keeps the contiguous block [r*nq/G, (r+1)*nq/G)
    ctx->nq_total = nq;
    ctx->q_begin = nq * ctx->rank / ctx->world;
    const int64_t q_end = nq * (ctx->rank + 1) / ctx->world;
    ctx->nq = q_end - ctx->q_begin;
    ctx->ntaxa = ntaxa;
    ctx->has_net = false;
    ctx->has_sampled = false;
    ctx->has_ci = false;
    cudaFree(ctx->d_ci_lo); cudaFree(ctx->d_ci_hi);
    ctx->d_ci_lo = ctx->d_ci_hi = nullptr;
    ctx->data_version++;
    cudaFree(ctx->d_taxa); cudaFree(ctx->d_obs); cudaFree(ctx->d_sampled); cudaFree(ctx->d_off); cudaFree(ctx->d_W);
    cudaFree(ctx->d_WA); cudaFree(ctx->d_wpart); cudaFree(ctx->d_negw3);
    ctx->d_WA = nullptr; ctx->d_wpart = nullptr; ctx->d_negw3 = nullptr;
    cudaFree(ctx->d_perm); cudaFree(ctx->d_qhdr); cudaFree(ctx->d_len4); cudaFree(ctx->d_iota); cudaFree(ctx->d_cls);
    cudaFree(ctx->d_cls_sorted);
    ctx->d_taxa = nullptr; ctx->d_obs = nullptr; ctx->d_sampled = nullptr; ctx->d_off = nullptr; ctx->d_W = nullptr;
    ctx->d_perm = nullptr; ctx->d_qhdr = nullptr; ctx->d_len4 = nullptr; ctx->d_iota = nullptr; ctx->d_cls = nullptr;
    ctx->d_cls_sorted = nullptr;
    const size_t n = (size_t)std::max<int64_t>(ctx->nq, 1);
    CK(cudaMalloc((void **)&ctx->d_taxa, n * 4 * sizeof(uint16_t)));
    CK(cudaMalloc((void **)&ctx->d_obs, n * 3 * sizeof(double)));
    CK(cudaMalloc((void **)&ctx->d_sampled, n));
    CK(cudaMalloc((void **)&ctx->d_off, (n + 1) * sizeof(uint32_t)));
    CK(cudaMalloc((void **)&ctx->d_W, n * 4 * sizeof(double)));
    CK(cudaMalloc((void **)&ctx->d_WA, n * sizeof(double2)));
    CK(cudaMalloc((void **)&ctx->d_wpart, 2 * ((n + 255) / 256) * sizeof(double)));
    CK(cudaMalloc((void **)&ctx->d_negw3, 2 * sizeof(double)));
    CK(cudaMalloc((void **)&ctx->d_perm, n * sizeof(uint32_t)));
    CK(cudaMalloc((void **)&ctx->d_qhdr, n));
    CK(cudaMalloc((void **)&ctx->d_len4, n * sizeof(uint32_t)));
    CK(cudaMalloc((void **)&ctx->d_iota, n * sizeof(uint32_t)));
    cudaFree(ctx->d_stage); cudaFree(ctx->d_hdr0);
    ctx->d_stage = nullptr; ctx->d_hdr0 = nullptr;
    CK(cudaMalloc((void **)&ctx->d_stage, n * SNAQ_STAGE_WORDS * sizeof(uint16_t)));
    CK(cudaMalloc((void **)&ctx->d_hdr0, n));
    cudaFree(ctx->d_flags); cudaFree(ctx->d_uidx);
    ctx->d_flags = ctx->d_uidx = nullptr;
    CK(cudaMalloc((void **)&ctx->d_flags, n * sizeof(uint32_t)));      // run table (k_eval_runs) and program deduplication
    CK(cudaMalloc((void **)&ctx->d_uidx, n * sizeof(uint32_t)));
    CK(cudaMalloc((void **)&ctx->d_cls, n * sizeof(uint32_t)));
    CK(cudaMalloc((void **)&ctx->d_cls_sorted, n * sizeof(uint32_t)));
    if (ctx->nq >= 0xFFFFFFFFLL) { ctx->err = "more than 2^32 quartets on one GPU: shard them (world_size>1)"; return SNAQ_EINVAL; }
    if (ctx->nq > 0) {
        CK(cudaMemcpyAsync(ctx->d_taxa, taxa + 4 * ctx->q_begin, (size_t)ctx->nq * 4 * sizeof(uint16_t), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(ctx->d_obs, obscf + 3 * ctx->q_begin, (size_t)ctx->nq * 3 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        ctx->counters[2] += ctx->nq * (4 * 2 + 3 * 8);
    }
    return SNAQ_OK;
}

int snaq_set_obscf(snaq_ctx *ctx, const double *obscf) {
    if (!ctx || !obscf) return SNAQ_EINVAL;
    if (!ctx->d_obs) { ctx->err = "snaq_set_data has not been called"; return SNAQ_ENODATA; }
    CK(cudaSetDevice(ctx->device));
    ctx->data_version++;
    for (int64_t i = 0; i < ctx->nq_total * 3; i++)
        if (!(obscf[i] >= 0.0 && obscf[i] <= 1.0)) { ctx->err = "obsCF must be between (0,1) in quartet " + std::to_string(i / 3); return SNAQ_EINVAL; }
    if (ctx->nq > 0) {
        CK(cudaMemcpyAsync(ctx->d_obs, obscf + 3 * ctx->q_begin, (size_t)ctx->nq * 3 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        ctx->counters[2] += ctx->nq * 24;
    }
    int rc = run_weights(ctx);
    if (rc) return rc;
    CK(cudaStreamSynchronize(ctx->stream));
    return SNAQ_OK;
}

int snaq_set_ci(snaq_ctx *ctx, const double *lo, const double *hi) {
    if (!ctx || !lo || !hi) return SNAQ_EINVAL;
    if (!ctx->d_obs) { ctx->err = "snaq_set_data has not been called"; return SNAQ_ENODATA; }
    CK(cudaSetDevice(ctx->device));
    for (int64_t q = 0; q < ctx->nq_total; q++) {
        double top = 0.0;
        for (int i = 0; i < 3; i++) {
            const double a = lo[3 * q + i], b = hi[3 * q + i];
            if (!(a >= 0.0 && a <= b && b <= 1.0)) { ctx->err = "credibility interval must satisfy 0 <= lo <= hi <= 1 in quartet " + std::to_string(q); return SNAQ_EINVAL; }
            top += b;
        }
        if (!(top > 0.0)) { ctx->err = "credibility intervals of quartet " + std::to_string(q) + " are all zero"; return SNAQ_EINVAL; }
    }
    const size_t n = (size_t)std::max<int64_t>(ctx->nq, 1) * 3;
    if (!ctx->d_ci_lo) CK(cudaMalloc((void **)&ctx->d_ci_lo, n * sizeof(double)));
    if (!ctx->d_ci_hi) CK(cudaMalloc((void **)&ctx->d_ci_hi, n * sizeof(double)));
    if (ctx->nq > 0) {
        CK(cudaMemcpyAsync(ctx->d_ci_lo, lo + 3 * ctx->q_begin, (size_t)ctx->nq * 3 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(ctx->d_ci_hi, hi + 3 * ctx->q_begin, (size_t)ctx->nq * 3 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        ctx->counters[2] += ctx->nq * 48;
    }
    ctx->has_ci = true;
    return SNAQ_OK;
}

int snaq_resample_obscf(snaq_ctx *ctx, uint64_t seed) {
    if (!ctx) return SNAQ_EINVAL;
    if (!ctx->has_ci) { ctx->err = "snaq_set_ci has not been called"; return SNAQ_ENODATA; }
    CK(cudaSetDevice(ctx->device));
    ctx->data_version++;
    if (ctx->nq > 0) {
        k_resample_ci<<<(unsigned)((ctx->nq + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_ci_lo, ctx->d_ci_hi, ctx->nq, ctx->q_begin,
                                                                                 (unsigned long long)seed, ctx->d_obs);
        CK(cudaGetLastError());
        ctx->counters[5] += 1;
    }
    int rc = run_weights(ctx);
    if (rc) return rc;
    CK(cudaStreamSynchronize(ctx->stream));
    return SNAQ_OK;
}

int snaq_get_obscf(snaq_ctx *ctx, double *obscf) {
    if (!ctx || !obscf) return SNAQ_EINVAL;
    if (!ctx->d_obs) { ctx->err = "snaq_set_data has not been called"; return SNAQ_ENODATA; }
    CK(cudaSetDevice(ctx->device));
    if (ctx->nq > 0) {
        CK(cudaMemcpyAsync(obscf + 3 * ctx->q_begin, ctx->d_obs, (size_t)ctx->nq * 3 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        ctx->counters[3] += ctx->nq * 24;
    }
    return SNAQ_OK;
}

int snaq_set_sampled(snaq_ctx *ctx, const uint8_t *mask) {
    if (!ctx) return SNAQ_EINVAL;
    if (!ctx->d_obs) { ctx->err = "snaq_set_data has not been called"; return SNAQ_ENODATA; }
    CK(cudaSetDevice(ctx->device));
    ctx->data_version++;
    ctx->has_sampled = (mask != nullptr);
    if (mask && ctx->nq > 0) {
        CK(cudaMemcpyAsync(ctx->d_sampled, mask + ctx->q_begin, (size_t)ctx->nq, cudaMemcpyHostToDevice, ctx->stream));
        ctx->counters[2] += ctx->nq;
    }
    int rc = run_weights(ctx);
    if (rc) return rc;
    CK(cudaStreamSynchronize(ctx->stream));
    return SNAQ_OK;
}

// SNAQ_BUILD_TIMING=1: host wall time of the stages of one table build (between the points where the host waits), to stderr
struct BuildClock {
    bool on;
    std::chrono::steady_clock::time_point t0, last;
    std::string line;
    BuildClock() : on(getenv("SNAQ_BUILD_TIMING") && atoi(getenv("SNAQ_BUILD_TIMING")) != 0) { t0 = last = std::chrono::steady_clock::now(); }
    void mark(snaq_ctx *ctx, const char *what);
    ~BuildClock() {
        if (on) fprintf(stderr, "[snaq build] %s total %.1f us\n", line.c_str(),
                        std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count());
    }
};

// Device side of snaq_set_network / snaq_patch_network: uploads the host tables ctx->H and (re)builds the class-sorted
// program table, the run table and the weights.  d_dirty (device, one byte per taxon) or nullptr: with it, only the quartets
// that contain a flagged taxon are classified again; everything downstream (sort, offsets, emission from the staged
// programs, runs, weights) is redone for the whole table, so the result is bit-identical to a build from scratch.
void BuildClock::mark(snaq_ctx *ctx, const char *what) {
    if (!on) return;
    cudaStreamSynchronize(ctx->stream);
    const auto now = std::chrono::steady_clock::now();
    char buf[64];
    snprintf(buf, sizeof buf, "%s %.1f | ", what, std::chrono::duration<double, std::micro>(now - last).count());
    line += buf;
    last = now;
}

static int build_table(snaq_ctx *ctx, const uint8_t *d_dirty) {
    const SnaqHostTables &H = ctx->H;
    BuildClock clk;
    int rc = ensure(ctx, ctx->d_blob, ctx->blob_cap, H.blob.size());
    if (rc) return rc;
    rc = ensure(ctx, ctx->d_xconst, ctx->xconst_cap, std::max<size_t>(H.xconst.size(), 1));
    if (rc) return rc;
    CK(cudaMemcpyAsync(ctx->d_blob, H.blob.data(), H.blob.size(), cudaMemcpyHostToDevice, ctx->stream));
    if (!H.xconst.empty())
        CK(cudaMemcpyAsync(ctx->d_xconst, H.xconst.data(), H.xconst.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    ctx->counters[2] += (int64_t)H.blob.size() + (int64_t)H.xconst.size() * 8;
    ctx->T = H.view(ctx->d_blob);
    if (ctx->nq > 0 && H.ntaxa > 0 && H.ntaxa <= SNAQ_LCA_MAX_TAXA) {
        const size_t n2 = (size_t)H.ntaxa * H.ntaxa;
        rc = ensure(ctx, ctx->d_lca, ctx->lca_cap, n2);
        if (rc) return rc;
        SnaqTables Tw = ctx->T;
        k_leaf_lca<<<(unsigned)((n2 + 255) / 256), 256, 0, ctx->stream>>>(Tw, ctx->d_lca);
        CK(cudaGetLastError());
        ctx->T.leaf_lca = ctx->d_lca;
    }
#ifdef SNAQ_BOUNDS
    {   // process-wide bound: the largest expanded vector of any live network (contexts run concurrently)
        static std::mutex mu;
        static int nf_max = 0;
        std::lock_guard<std::mutex> lock(mu);
        if (H.n_full > nf_max) { nf_max = H.n_full; CK(cudaDeviceSynchronize()); CK(cudaMemcpyToSymbol(g_dbg_nfull, &nf_max, sizeof(int))); }
    }
#endif
    clk.mark(ctx, "upload+lca");
    const int64_t nq = ctx->nq;
    ctx->prog_words = 0;
    ctx->maxlen = 1;
    ctx->nA1 = ctx->nA2 = ctx->nA = ctx->nB = ctx->nC = 0;
    if (nq > 0) {
        // pass 1: program length (4-word chunks) and class of every quartet, original order
        CK(cudaMemsetAsync(ctx->d_total, 0, 8 * sizeof(unsigned long long), ctx->stream));
        CK(cudaMemsetAsync(ctx->d_status, 0, 4 * sizeof(uint32_t), ctx->stream));
        const int64_t nb = (nq + 127) / 128;
        const uint32_t *d_list = nullptr, *d_nlist = nullptr;
        if (d_dirty) {
            // (d_perm is rebuilt by the sort below: until then it holds the list; d_total[7] its length)
            d_list = ctx->d_perm; d_nlist = reinterpret_cast<const uint32_t *>(ctx->d_total + 7);
            k_dirty_list<<<(unsigned)((nq + 1023) / 1024), 256, 0, ctx->stream>>>(ctx->d_taxa, nq, d_dirty, ctx->d_len4, ctx->d_cls, ctx->d_perm,
                                                                                reinterpret_cast<uint32_t *>(ctx->d_total + 7), ctx->d_total);
            CK(cudaGetLastError());
        }
        k_build_count<<<(unsigned)nb, 128, 0, ctx->stream>>>(ctx->T, ctx->d_taxa, nq, ctx->dedup ? 1 : 0, ctx->d_len4, ctx->d_cls, ctx->d_iota,
                                                            ctx->d_total, ctx->d_status, d_list, d_nlist, ctx->d_stage, ctx->d_hdr0);
        CK(cudaGetLastError());
        unsigned long long stt[7] = {0, 0, 0, 0, 0, 0, 0};
        uint32_t st[4];
        CK(cudaMemcpyAsync(stt, ctx->d_total, sizeof(stt), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(st, ctx->d_status, sizeof(st), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        clk.mark(ctx, "count");
        if (st[0]) { ctx->err = build_err_text(st[0]); return SNAQ_ETOPOLOGY; }
        if (stt[0] >= 0xFFFFFFFFull) { ctx->err = "descriptor table exceeds 2^32 chunks on one GPU: shard the quartets (world_size>1)"; return SNAQ_EINVAL; }
        ctx->prog_words = stt[0] * 4;
        ctx->maxlen = (uint32_t)stt[1];
        ctx->nA1 = (int64_t)stt[2]; ctx->nA2 = (int64_t)stt[3];
        ctx->nA = (int64_t)(stt[2] + stt[3] + stt[4]); ctx->nB = (int64_t)stt[5]; ctx->nC = (int64_t)stt[6];
        // stable sort of the quartet indices by (class, term hash) (31-bit keys): perm = sorted position -> original quartet
        size_t tmp = 0, tmp2 = 0;
        CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp, ctx->d_cls, ctx->d_cls_sorted, ctx->d_iota, ctx->d_perm, nq, 0, 31, ctx->stream));
        uint32_t *d_len_sorted = ctx->d_off + 1;
        CK(cub::DeviceScan::InclusiveSum(nullptr, tmp2, d_len_sorted, d_len_sorted, nq, ctx->stream));
        tmp = std::max(tmp, tmp2);
        if (tmp > ctx->scan_tmp_cap) {
            cudaFree(ctx->d_scan_tmp);
            ctx->d_scan_tmp = nullptr;
            CK(cudaMalloc(&ctx->d_scan_tmp, tmp));
            ctx->scan_tmp_cap = tmp;
        }
        size_t t1 = ctx->scan_tmp_cap;
        CK(cub::DeviceRadixSort::SortPairs(ctx->d_scan_tmp, t1, ctx->d_cls, ctx->d_cls_sorted, ctx->d_iota, ctx->d_perm, nq, 0, 31, ctx->stream));
        // lengths in sorted order, scanned in place => d_off[i+1] = end of program i (chunks)
        CK(cudaMemsetAsync(ctx->d_off, 0, sizeof(uint32_t), ctx->stream));
        const int64_t pad_at = (ctx->nA1 & 1) ? ctx->nA1 - 1 : -1;
        if (pad_at >= 0) ctx->prog_words += 4;
        k_gather_len<<<(unsigned)((nq + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_perm, ctx->d_len4, nq, pad_at, d_len_sorted);
        CK(cudaGetLastError());
        t1 = ctx->scan_tmp_cap;
        CK(cub::DeviceScan::InclusiveSum(ctx->d_scan_tmp, t1, d_len_sorted, d_len_sorted, nq, ctx->stream));
        rc = ensure(ctx, ctx->d_prog, ctx->prog_cap, (size_t)ctx->prog_words + 8);
        if (rc) return rc;
        k_build_emit<<<(unsigned)nb, 128, 0, ctx->stream>>>(ctx->T, ctx->d_taxa, nq, ctx->d_perm, ctx->d_off, ctx->d_prog, ctx->d_qhdr,
                                                           ctx->d_status, ctx->d_len4, ctx->d_stage, ctx->d_hdr0);
        CK(cudaGetLastError());
        ctx->counters[5] += 5;
        ctx->nU = 0;
        clk.mark(ctx, "sort+emit");
        rc = build_dedup(ctx, pad_at);                     // run table: one program per run of identical programs
        if (rc) return rc;
        clk.mark(ctx, "runs");
        rc = build_run_tiles(ctx);                         // tiles of the run table (k_eval_runs, run-summed weights)
        if (rc) return rc;
        clk.mark(ctx, "tiles");
        ctx->has_net = true;
        rc = run_weights(ctx);
        if (rc) { ctx->has_net = false; return rc; }
        clk.mark(ctx, "weights");
        CK(cudaMemcpyAsync(st, ctx->d_status, sizeof(st), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        if (st[0]) { ctx->has_net = false; ctx->err = build_err_text(st[0]); return SNAQ_ETOPOLOGY; }
        ctx->counters[3] += 40 + 16 + 16;
    }
    ctx->counters[4] = (int64_t)ctx->prog_words;
    {   // bytes one single-call evaluation (k_eval_direct, P <= 4) streams from HBM
        const int64_t nG = ctx->nq - ctx->nA1 - ctx->nA2;
        const int64_t chunksG = (int64_t)(ctx->prog_words / 4) - ctx->nA1 - 2 * ctx->nA2;
        ctx->counters[6] = ctx->nA1 * 24 + ctx->nA2 * 32 + nG * (32 + 4 + 1) + chunksG * 8;
        ctx->counters[7] = ctx->nU;
    }
    ctx->has_net = true;
    return SNAQ_OK;
}

int snaq_set_network(snaq_ctx *ctx, const snaq_flatnet *net) {
    if (!ctx) return SNAQ_EINVAL;
    if (!ctx->d_taxa) { ctx->err = "snaq_set_data has not been called"; return SNAQ_ENODATA; }
    CK(cudaSetDevice(ctx->device));
    ctx->has_net = false;
    int rc = snaq_build_tables(net, ctx->ntaxa, ctx->H, ctx->err);
    if (rc) return rc;
    nvtx_range r("snaq_set_network");
    return build_table(ctx, nullptr);
}

int snaq_patch_network(snaq_ctx *ctx, const snaq_flatnet *net, const int32_t *touched_taxa, int32_t n_touched) {
    if (!ctx) return SNAQ_EINVAL;
    if (!ctx->d_taxa) { ctx->err = "snaq_set_data has not been called"; return SNAQ_ENODATA; }
    if (n_touched > 0 && !touched_taxa) { ctx->err = "bad arguments to snaq_patch_network"; return SNAQ_EINVAL; }
    CK(cudaSetDevice(ctx->device));
    const bool had = ctx->has_net;
    ctx->has_net = false;
    // the new tables are rooted where the current ones are, so that taxa away from the edit keep their root paths
    SnaqHostTables Hn;
    int rc = snaq_build_tables(net, ctx->ntaxa, Hn, ctx->err, had ? ctx->H.root : -1);
    if (rc) return rc;
    std::vector<uint8_t> touched;
    bool partial = had && n_touched >= 0 && snaq_tables_touched(ctx->H, Hn, touched);
    if (partial) {
        for (int32_t k = 0; k < n_touched; k++) {                 // the caller's list is advisory: added to what the tables show
            if (touched_taxa[k] < 0 || touched_taxa[k] >= ctx->ntaxa) { ctx->err = "touched taxon id out of range"; return SNAQ_EINVAL; }
            touched[(size_t)touched_taxa[k]] = 1;
        }
        size_t nd = 0;
        for (uint8_t t : touched) nd += t;
        ctx->patch_touched = (int64_t)nd;
        if (nd * 2 > touched.size()) partial = false;             // most quartets would be re-described anyway
    }
    if (!partial) ctx->patch_touched = -1;
    ctx->H = std::move(Hn);
    nvtx_range r("snaq_patch_network");
    if (partial) {
        rc = ensure(ctx, ctx->d_dirty, ctx->dirty_cap, touched.size());
        if (rc) return rc;
        CK(cudaMemcpyAsync(ctx->d_dirty, touched.data(), touched.size(), cudaMemcpyHostToDevice, ctx->stream));
    }
    return build_table(ctx, partial ? ctx->d_dirty : nullptr);
}

int snaq_debug_bounds(snaq_ctx *ctx, int64_t *checks, int64_t *violations) {
    if (!ctx || !checks || !violations) return SNAQ_EINVAL;
#ifdef SNAQ_BOUNDS
    CK(cudaSetDevice(ctx->device));
    CK(cudaDeviceSynchronize());
    unsigned long long c = 0, b = 0;
    CK(cudaMemcpyFromSymbol(&c, g_bounds_checks, sizeof(c)));
    CK(cudaMemcpyFromSymbol(&b, g_bounds_bad, sizeof(b)));
    *checks = (int64_t)c; *violations = (int64_t)b;
#else
    *checks = -1; *violations = -1;            // not a bounds-checked build
#endif
    return SNAQ_OK;
}

int snaq_patch_info(snaq_ctx *ctx, int64_t *touched_taxa) {
    if (!ctx || !touched_taxa) return SNAQ_EINVAL;
    *touched_taxa = ctx->patch_touched;
    return SNAQ_OK;
}

int snaq_table_checksum(snaq_ctx *ctx, uint64_t out[4]) {
    if (!ctx || !out) return SNAQ_EINVAL;
    if (!ctx->has_net) { ctx->err = "snaq_set_network has not been called"; return SNAQ_ENODATA; }
    CK(cudaSetDevice(ctx->device));
    out[0] = out[1] = out[2] = 0;
    out[3] = ((uint64_t)ctx->nU << 32) | (uint64_t)ctx->rn_ntiles;
    if (ctx->nq == 0) return SNAQ_OK;
    unsigned long long *d = nullptr;
    CK(cudaMalloc((void **)&d, 3 * sizeof(unsigned long long)));
    cudaError_t e = cudaMemsetAsync(d, 0, 3 * sizeof(unsigned long long), ctx->stream);
    if (e == cudaSuccess) {
        k_checksum<<<(unsigned)((ctx->nq + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_perm, ctx->d_qhdr, ctx->d_off,
            reinterpret_cast<const unsigned long long *>(ctx->d_prog), reinterpret_cast<const unsigned long long *>(ctx->d_W), ctx->nq,
            (int64_t)(ctx->prog_words / 4), d);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d);
    if (e != cudaSuccess) { ctx->err = cudaGetErrorString(e); return SNAQ_ECUDA; }
    return SNAQ_OK;
}

int snaq_eval_topologies(snaq_ctx *ctx, int32_t C, const snaq_flatnet *nets, double *out) {
    if (!ctx) return SNAQ_EINVAL;
    if (C < 0 || (C > 0 && (!nets || !out))) { ctx->err = "bad arguments to snaq_eval_topologies"; return SNAQ_EINVAL; }
    for (int32_t c = 0; c < C; c++) {
        int rc = snaq_set_network(ctx, &nets[c]);
        if (rc) { ctx->err = "topology " + std::to_string(c + 1) + ": " + ctx->err; return rc; }
        const int np = ctx->H.nparams;
        rc = snaq_eval(ctx, 1, np, np > 0 ? ctx->H.x0.data() : &out[c], &out[c]);    // np == 0: X is not read
        if (rc) { ctx->err = "topology " + std::to_string(c + 1) + ": " + ctx->err; return rc; }
    }
    return SNAQ_OK;
}

// Brings worker w's data up to date with ctx's: a device-to-device copy of the three data arrays (allocation through
// snaq_set_data only when the shape changed).
static int sync_worker(snaq_ctx *ctx, size_t w) {
    snaq_ctx *wk = ctx->workers[w];
    if (ctx->worker_version[w] == ctx->data_version) return SNAQ_OK;
    if (wk->nq != ctx->nq || wk->ntaxa != ctx->ntaxa || !wk->d_taxa) {
        std::vector<uint16_t> taxa((size_t)ctx->nq * 4);
        std::vector<double> obs((size_t)ctx->nq * 3);
        if (ctx->nq > 0) {
            CK(cudaMemcpy(taxa.data(), ctx->d_taxa, taxa.size() * sizeof(uint16_t), cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(obs.data(), ctx->d_obs, obs.size() * sizeof(double), cudaMemcpyDeviceToHost));
        }
        int rc = snaq_set_data(wk, ctx->ntaxa, ctx->nq, taxa.data(), obs.data());
        if (rc) { ctx->err = "worker context: " + wk->err; return rc; }
    } else if (ctx->nq > 0) {
        CK(cudaMemcpyAsync(wk->d_obs, ctx->d_obs, (size_t)ctx->nq * 3 * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    }
    wk->has_sampled = ctx->has_sampled;
    if (ctx->has_sampled && ctx->nq > 0)
        CK(cudaMemcpyAsync(wk->d_sampled, ctx->d_sampled, (size_t)ctx->nq, cudaMemcpyDeviceToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    wk->has_net = false;                                     // the weights of its table are rebuilt by the next snaq_set_network
    wk->data_version++;
    ctx->worker_version[w] = ctx->data_version;
    return SNAQ_OK;
}

int snaq_optbl_topologies(snaq_ctx *ctx, int32_t C, const snaq_flatnet *nets, const snaq_optbl_opts *opts, int32_t workers,
                          double *xmin, int32_t ldx, int32_t *nparams, snaq_optbl_result *results) {
    if (!ctx) return SNAQ_EINVAL;
    if (C < 0 || workers < 0 || workers > 32 || (C > 0 && (!nets || !xmin || !results || ldx < 0))) {
        ctx->err = "bad arguments to snaq_optbl_topologies";
        return SNAQ_EINVAL;
    }
    if (!ctx->d_taxa) { ctx->err = "snaq_set_data has not been called"; return SNAQ_ENODATA; }
    if (ctx->world > 1) { ctx->err = "snaq_optbl_topologies is not available in quartet-sharded mode"; return SNAQ_EINVAL; }
    if (C == 0) return SNAQ_OK;
    CK(cudaSetDevice(ctx->device));
    const size_t W = (size_t)std::min<int32_t>(workers > 0 ? workers : 8, C);
    while (ctx->workers.size() < W) {
        snaq_opts o{ctx->device, 0, 1, ctx->dedup ? SNAQ_OPT_DEDUP : SNAQ_OPT_PER_QUARTET};
        snaq_ctx *wk = nullptr;
        int rc = snaq_create(&o, &wk);
        if (rc) { ctx->err = std::string("worker context: ") + snaq_last_error(nullptr); return rc; }
        wk->reuse = ctx->reuse; wk->eval_chunks = ctx->eval_chunks; wk->lanes_persistent = ctx->lanes_persistent;
        wk->lanes_chunk = ctx->lanes_chunk; wk->lanes_warps = ctx->lanes_warps;
        ctx->workers.push_back(wk);
        ctx->worker_version.push_back(0);
    }
    for (size_t w = 0; w < W; w++) {
        int rc = sync_worker(ctx, w);
        if (rc) return rc;
    }
    std::atomic<int32_t> next{0};
    std::vector<int> rcs((size_t)C, SNAQ_OK);
    std::vector<std::string> msgs((size_t)C);
    auto body = [&](size_t w) {
        snaq_ctx *wk = ctx->workers[w];
        for (;;) {
            const int32_t c = next.fetch_add(1);
            if (c >= C) break;
            snaq_optbl_result r{};
            int32_t np = 0;
            int rc = snaq_set_network(wk, &nets[c]);
            if (!rc) rc = snaq_num_params(wk, &np, nullptr, nullptr, nullptr);
            if (!rc && np > ldx) { wk->err = "ldx is smaller than the number of parameters"; rc = SNAQ_EINVAL; }
            if (!rc) {
                double *x = xmin + (size_t)c * ldx;
                rc = snaq_get_params(wk, x, nullptr, nullptr);
                if (!rc) rc = snaq_optbl(wk, opts, x, &r);
            }
            if (nparams) nparams[c] = np;
            if (rc) { r = snaq_optbl_result{}; r.status = -rc; msgs[c] = wk->err; }
            rcs[c] = rc;
            results[c] = r;
        }
    };
    std::vector<std::thread> pool;
    for (size_t w = 1; w < W; w++) pool.emplace_back(body, w);
    body(0);
    for (auto &t : pool) t.join();
    for (size_t w = 0; w < W; w++)
        for (int k : {0, 1, 2, 3, 5}) { ctx->counters[k] += ctx->workers[w]->counters[k]; ctx->workers[w]->counters[k] = 0; }   // the cumulative ones
    for (int32_t c = 0; c < C; c++)
        if (rcs[c]) { ctx->err = "topology " + std::to_string(c + 1) + ": " + msgs[c]; return rcs[c]; }
    return SNAQ_OK;
}

int snaq_num_params(snaq_ctx *ctx, int32_t *nparams, int32_t *n_gamma, int32_t *n_t, int32_t *n_gammaz) {
    if (!ctx) return SNAQ_EINVAL;
    if (!ctx->has_net) { ctx->err = "snaq_set_network has not been called"; return SNAQ_ENODATA; }
    if (nparams) *nparams = ctx->H.nparams;
    if (n_gamma) *n_gamma = ctx->H.nh;
    if (n_t) *n_t = ctx->H.nt;
    if (n_gammaz) *n_gammaz = ctx->H.nhz;
    return SNAQ_OK;
}

int snaq_get_params(snaq_ctx *ctx, double *x, int32_t *numht, double *upper) {
    if (!ctx) return SNAQ_EINVAL;
    if (!ctx->has_net) { ctx->err = "snaq_set_network has not been called"; return SNAQ_ENODATA; }
    const SnaqHostTables &H = ctx->H;
    if (x) std::memcpy(x, H.x0.data(), sizeof(double) * H.nparams);
    if (numht) std::memcpy(numht, H.numht.data(), sizeof(int32_t) * H.nparams);
    if (upper) std::memcpy(upper, H.upper.data(), sizeof(double) * H.nparams);
    return SNAQ_OK;
}

static int check_x_host(snaq_ctx *ctx, int P, const double *X) {
    const SnaqHostTables &H = ctx->H;
    for (int p = 0; p < P; p++)
        for (int s = 0; s < H.nparams; s++) {
            const double v = X[(size_t)p * H.nparams + s];
            const bool isT = s >= H.nh && s < H.nh + H.nt;
            if (!(v >= 0.0) || (!isT && v > 1.0)) {
                char buf[256];
                if (isT) snprintf(buf, sizeof buf, "length has to be nonnegative: %g (parameter %d of vector %d)", v, s + 1, p + 1);
                else snprintf(buf, sizeof buf, "new %s value should be between 0,1: %g. (parameter %d of vector %d)", s < H.nh ? "gamma" : "gammaz", v, s + 1, p + 1);
                ctx->err = buf;
                return SNAQ_ERANGE;
            }
        }
    return SNAQ_OK;
}

int snaq_eval_device(snaq_ctx *ctx, int32_t P, int32_t nparams, const double *dX, double *dout, void *stream) {
    if (!ctx) return SNAQ_EINVAL;
    if (!ctx->has_net) { ctx->err = "snaq_set_network has not been called"; return SNAQ_ENODATA; }
    if (P <= 0 || nparams != ctx->H.nparams || !dX || !dout) { ctx->err = "bad arguments to snaq_eval_device (x has wrong length?)"; return SNAQ_EINVAL; }
    CK(cudaSetDevice(ctx->device));
    return eval_device(ctx, P, dX, dout, stream ? (cudaStream_t)stream : ctx->stream, nullptr, nullptr, nullptr);
}

// status word of the evaluation kernels since the last call; clears it
int snaq_check_status(snaq_ctx *ctx) {
    if (!ctx) return SNAQ_EINVAL;
    CK(cudaSetDevice(ctx->device));
    uint32_t st = 0;
    CK(cudaMemcpy(&st, ctx->d_status + 1, sizeof(st), cudaMemcpyDeviceToHost));
    if (st) CK(cudaMemset(ctx->d_status + 1, 0, sizeof(uint32_t)));
    return status_to_error(ctx, st);
}

int snaq_eval(snaq_ctx *ctx, int32_t P, int32_t nparams, const double *X, double *out) {
    if (!ctx) return SNAQ_EINVAL;
    if (!ctx->has_net) { ctx->err = "snaq_set_network has not been called"; return SNAQ_ENODATA; }
    if (P <= 0 || !X || !out) { ctx->err = "bad arguments to snaq_eval"; return SNAQ_EINVAL; }
    if (nparams != ctx->H.nparams) {
        ctx->err = "ht(net) (length " + std::to_string(ctx->H.nparams) + ") and vector x (length " + std::to_string(nparams) + ") need to have same length";
        return SNAQ_EINVAL;
    }
    CK(cudaSetDevice(ctx->device));
    const size_t nx = (size_t)P * std::max(nparams, 1);
    int rc = ensure(ctx, ctx->d_X, ctx->X_cap, nx);
    if (rc) return rc;
    rc = ensure(ctx, ctx->d_out, ctx->out_cap, (size_t)P + 1);
    if (rc) return rc;
    rc = ensure_pin(ctx, nx + P + 2);
    if (rc) return rc;
    if (P < LANES_MIN_P && ctx->nq > 0 && (size_t)std::min(P, 4) * nparams <= XARG_MAX && (!ctx->comm || fused_ok(ctx))) {
        // small batch (a single objective call of the optimiser): the candidates travel as kernel arguments and the last
        // block writes the results and the status word straight into page-locked host memory: no copy calls at all
        double *h_out = ctx->h_pin;
        uint32_t *h_st = reinterpret_cast<uint32_t *>(h_out + P);
        *h_st = 0;
        ctx->xarg_src = X;
        ctx->xarg_base = ctx->d_X;
        ctx->status_out = h_st;
        rc = eval_device(ctx, P, ctx->d_X, h_out, ctx->stream, nullptr, nullptr, nullptr);
        ctx->xarg_src = nullptr;
        ctx->status_out = nullptr;
        if (rc) return rc;
        CK(wait_done(ctx->stream, h_st));
        ctx->counters[2] += (int64_t)sizeof(double) * P * nparams;
        ctx->counters[3] += (int64_t)sizeof(double) * P + 4;
        if (*h_st) {
            const uint32_t st = *h_st;
            CK(cudaMemsetAsync(ctx->d_status + 1, 0, sizeof(uint32_t), ctx->stream));
            if (st & SNAQ_S_RANGE) { rc = check_x_host(ctx, P, X); if (rc) return rc; }
            return status_to_error(ctx, st);
        }
        std::memcpy(out, h_out, sizeof(double) * P);
        return SNAQ_OK;
    }
    // the bounds of update! (src/snaq_optimization.jl:213,221,228) are checked on the device; only when the device
    // reports a violation is the batch scanned on the host for the message.  A caller buffer that is already
    // page-locked is copied from directly; otherwise it is staged through the context's pinned buffer.
    cudaPointerAttributes attr;
    const bool user_pinned = cudaPointerGetAttributes(&attr, X) == cudaSuccess && attr.type == cudaMemoryTypeHost;
    cudaGetLastError();
    const double *src = user_pinned ? X : ctx->h_pin;
    // large batches: the copy of chunk k+1 (copy stream) overlaps the evaluation of chunk k (compute stream); a pageable
    // caller buffer is staged chunk by chunk, so that the host memcpy of chunk k+1 also overlaps the device work of chunk k
    const int nchunk = (P >= 2048 && nparams > 0) ? ctx->eval_chunks : 1;   // (measured: two chunks are best for pinned and pageable alike)
    if (nchunk == 1 && !user_pinned) std::memcpy(ctx->h_pin, X, sizeof(double) * (size_t)P * nparams);
    if (nchunk > 1) {
        if (!ctx->copy_stream) {
            CK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
            for (int k = 0; k < 8; k++) CK(cudaEventCreateWithFlags(&ctx->copy_ev[k], cudaEventDisableTiming));
        }
        const int Pc = ((P + nchunk - 1) / nchunk + 31) & ~31;
        int k = 0;
        for (int p0 = 0; p0 < P; p0 += Pc, k++) {
            const int pn = std::min(Pc, P - p0);
            if (!user_pinned) std::memcpy(ctx->h_pin + (size_t)p0 * nparams, X + (size_t)p0 * nparams, sizeof(double) * (size_t)pn * nparams);
            CK(cudaMemcpyAsync(ctx->d_X + (size_t)p0 * nparams, src + (size_t)p0 * nparams, sizeof(double) * (size_t)pn * nparams,
                               cudaMemcpyHostToDevice, ctx->copy_stream));
            CK(cudaEventRecord(ctx->copy_ev[k], ctx->copy_stream));
            CK(cudaStreamWaitEvent(ctx->stream, ctx->copy_ev[k], 0));
            rc = eval_device(ctx, pn, ctx->d_X + (size_t)p0 * nparams, ctx->d_out + p0, ctx->stream, nullptr, nullptr, nullptr);
            if (rc) return rc;
        }
    } else {
        CK(cudaMemcpyAsync(ctx->d_X, src, sizeof(double) * (size_t)P * nparams, cudaMemcpyHostToDevice, ctx->stream));
        rc = eval_device(ctx, P, ctx->d_X, ctx->d_out, ctx->stream, nullptr, nullptr, nullptr);
        if (rc) return rc;
    }
    double *h_out = ctx->h_pin + nx;
    uint32_t *h_st = reinterpret_cast<uint32_t *>(h_out + P);
    CK(cudaMemcpyAsync(h_out, ctx->d_out, sizeof(double) * P, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(h_st, ctx->d_status + 1, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->counters[2] += (int64_t)sizeof(double) * P * nparams;
    ctx->counters[3] += (int64_t)sizeof(double) * P + 4;
    if (*h_st) {
        CK(cudaMemsetAsync(ctx->d_status + 1, 0, sizeof(uint32_t), ctx->stream));
        if (*h_st & SNAQ_S_RANGE) { rc = check_x_host(ctx, P, X); if (rc) return rc; }
        return status_to_error(ctx, *h_st);
    }
    std::memcpy(out, h_out, sizeof(double) * P);
    return SNAQ_OK;
}

int snaq_eval_grad(snaq_ctx *ctx, int32_t nparams, const double *x, double *f, double *grad, double *hdiag) {
    if (!ctx) return SNAQ_EINVAL;
    if (!ctx->has_net) { ctx->err = "snaq_set_network has not been called"; return SNAQ_ENODATA; }
    if (!x || !f || !grad || nparams != ctx->H.nparams) { ctx->err = "bad arguments to snaq_eval_grad (x has wrong length?)"; return SNAQ_EINVAL; }
    if (ctx->world > 1 && !ctx->comm) {
        ctx->err = "snaq_eval_grad / snaq_optbl on a quartet-sharded context need a communicator (snaq_comm_init): a rank's partial "
                   "sums are not the objective";
        return SNAQ_EINVAL;
    }
    int rc = check_x_host(ctx, 1, x);
    if (rc) return rc;
    CK(cudaSetDevice(ctx->device));
    const SnaqHostTables &H = ctx->H;
    const int n_full = H.n_full;
    for (int i = 0; i < nparams; i++) grad[i] = 0.0;
    if (hdiag) for (int i = 0; i < nparams; i++) hdiag[i] = 0.0;
    if (ctx->nq == 0 && !ctx->comm) { *f = 0.0; return SNAQ_OK; }
    const int nrows = hdiag ? 2 * n_full : n_full;
    const size_t smem_base = 640 + 8 * QD_SPAN * 8 + (size_t)n_full * 8 + (size_t)nrows * 12 + 8 + (size_t)nparams * 8;   // xf, 96-bit bins, x
    if (smem_base - 128 > 100 * 1024) { ctx->err = "network too large for the shared-memory tile of k_eval_grad"; return SNAQ_EINVAL; }
    EvalTable tb = eval_table(ctx, false, true);
    const bool compressed = use_compressed(ctx, false, true);
    if (compressed) tb.nA1 = tb.nA2 = 0;                    // see launch_direct
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((tb.nq + 255) / 256, (int64_t)ctx->sm_count * 3));
    // the streaming kernel (class-A ring, program memo, grid-wide integer bins) pays from about 65,000 class-A quartets on
    // (measured: 50 taxa and up; at 30 taxa the two are equal, below the small kernel wins by 4-8 us); else k_eval_grad_small
    const int64_t tilesA = (tb.nA1 + QD_TQ - 1) / QD_TQ + (tb.nA2 + QD_TQ - 1) / QD_TQ;
    const bool use_ring = tilesA >= 1024 && smem_base + 8 * QD_NST * QD_STAGE_BYTES <= 110 * 1024;   // (a very large network: no room for the ring)
    const bool agg_on = !compressed && tb.nq >= 65536;
    // (k_eval_grad_small lays its shared memory out without the barrier area: 128 bytes less)
    const size_t smem = use_ring ? smem_base + 8 * QD_NST * QD_STAGE_BYTES : smem_base - 128;
    if (!use_ring) {
        rc = ensure(ctx, ctx->d_gpart, ctx->gpart_cap, (size_t)grid * nrows);
        if (rc) return rc;
    }
    const int kflags = agg_on ? 0 : 2;
    rc = ensure(ctx, ctx->d_X, ctx->X_cap, (size_t)std::max(nparams, 1));
    if (rc) return rc;
    rc = ensure(ctx, ctx->d_partial, ctx->partial_cap, (size_t)grid);
    if (rc) return rc;
    const size_t gb_need = (size_t)4 * n_full;                            // [2][2 n_full] grid bins (sized for the curvature variant too)
    if (gb_need > ctx->gbins_cap || !ctx->d_gbins) {                      // zero once, the kernel re-zeroes its bins
        rc = ensure(ctx, ctx->d_gbins, ctx->gbins_cap, gb_need);
        if (rc) return rc;
        CK(cudaMemsetAsync(ctx->d_gbins, 0, ctx->gbins_cap * sizeof(unsigned long long), ctx->stream));
    }
    rc = ensure(ctx, ctx->d_gout, ctx->gout_cap, (size_t)nrows + 1);
    if (rc) return rc;
    rc = ensure_pin(ctx, (size_t)nparams + nrows + 4);
    if (rc) return rc;
    // x travels as a kernel argument when it fits; the value, the gradient rows and the status word are written by the last
    // block straight into page-locked host memory: the call is one launch and one stream synchronisation
    const bool by_arg = nparams <= XARG_MAX;
    double *h_g = ctx->h_pin + nparams;
    uint32_t *h_st = reinterpret_cast<uint32_t *>(h_g + nrows + 1);
    *h_st = 0;
    // quartet-sharded: the kernel leaves this rank's value and rows in device memory, they are summed over the ranks
    // (ncclAllReduce on the same stream) and only then copied to the host: every rank gets the same totals
    const bool sharded = ctx->comm != nullptr;
    double *o_f = h_g;
    uint32_t *o_st = h_st;
    if (sharded) {
        rc = ensure(ctx, ctx->d_red, ctx->red_cap, (size_t)nrows + 1);
        if (rc) return rc;
        o_f = ctx->d_red;
        o_st = nullptr;
        if (ctx->nq == 0) CK(cudaMemsetAsync(ctx->d_red, 0, sizeof(double) * ((size_t)nrows + 1), ctx->stream));
    }
    if (by_arg) std::memcpy(ctx->xarg.v, x, sizeof(double) * nparams);
    else {
        std::memcpy(ctx->h_pin, x, sizeof(double) * nparams);
        CK(cudaMemcpyAsync(ctx->d_X, ctx->h_pin, sizeof(double) * nparams, cudaMemcpyHostToDevice, ctx->stream));
    }
    if (ctx->d_trace) ctx->trace_blocks = use_ring ? grid : 0;
#define LAUNCH_GRAD(WH)                                                                                                        \
    do {                                                                                                                       \
        if (use_ring) {                                                                                                        \
            CK(raise_smem_limit(k_eval_grad<WH>, smem));                                                                         \
            k_eval_grad<WH><<<grid, 256, smem, ctx->stream>>>(                                                                  \
                ctx->T, tb.off, reinterpret_cast<const uint2 *>(tb.prog), tb.qhdr, tb.WA,                                       \
                reinterpret_cast<const double4 *>(tb.W), (uint32_t)tb.nA1, (uint32_t)tb.nA2,                                    \
                (uint32_t)(tb.nA1 + (tb.nA1 & 1)), (uint32_t)tb.nq, ctx->d_X, ctx->d_xconst, H.nh, H.nt, ctx->d_partial,        \
                o_f, tb.negw3, ctx->d_ticket, ctx->d_gbins, o_f + 1, ctx->d_status + 1, o_st, (by_arg ? 1 : 0) | kflags,        \
                tb.pen, ctx->d_trace, ctx->xarg);                                                                               \
        } else {                                                                                                               \
            CK(raise_smem_limit(k_eval_grad_small<WH>, smem));                                                                   \
            k_eval_grad_small<WH><<<grid, 256, smem, ctx->stream>>>(                                                            \
                ctx->T, tb.off, reinterpret_cast<const uint2 *>(tb.prog), tb.qhdr, tb.WA,                                       \
                reinterpret_cast<const double4 *>(tb.W), (uint32_t)tb.nA1, (uint32_t)tb.nA2,                                    \
                (uint32_t)(tb.nA1 + (tb.nA1 & 1)), (uint32_t)tb.nq, ctx->d_X, ctx->d_xconst, H.nh, H.nt, ctx->d_partial,        \
                o_f, tb.negw3, ctx->d_ticket, ctx->d_gpart, o_f + 1, ctx->d_status + 1, o_st, by_arg ? 1 : 0, tb.pen, ctx->xarg); \
        }                                                                                                                      \
    } while (0)
    if (ctx->nq > 0) { if (hdiag) LAUNCH_GRAD(true); else LAUNCH_GRAD(false); }
#undef LAUNCH_GRAD
    CK(cudaGetLastError());
    if (sharded) {
        rc = comm_allreduce(ctx, ctx->d_red, (size_t)nrows + 1, ctx->stream);
        if (rc) return rc;
        CK(cudaMemcpyAsync(h_g, ctx->d_red, sizeof(double) * ((size_t)nrows + 1), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(h_st, ctx->d_status + 1, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    } else CK(wait_done(ctx->stream, h_st));
    ctx->counters[0] += 1;
    ctx->counters[1] += ctx->nq;
    ctx->counters[2] += 8 * nparams;
    ctx->counters[3] += 8 * (nrows + 1) + 4;
    if (*h_st) {
        CK(cudaMemsetAsync(ctx->d_status + 1, 0, sizeof(uint32_t), ctx->stream));
        return status_to_error(ctx, *h_st);
    }
    *f = h_g[0];
    // chain rule through the expansion x -> xf, on the host tables
    const SnaqTables Th = H.view(H.blob.data());
    static const double tab[SNAQ_MATH_DOUBLES] = SNAQ_EXP_TAB_INIT_LOGS;
    const SnaqMath Mh{tab, tab + 16, tab + 32};
    snaq_grad_backprop(Th, Mh, x, H.xconst.data(), h_g + 1, grad);
    if (hdiag) snaq_grad_backprop(Th, Mh, x, H.xconst.data(), h_g + 1 + n_full, hdiag, true);
    return SNAQ_OK;
}

int snaq_update_bl(snaq_ctx *ctx, int64_t *nquartets, double *edgeL) {
    if (!ctx) return SNAQ_EINVAL;
    if (!ctx->has_net) { ctx->err = "snaq_set_network has not been called"; return SNAQ_ENODATA; }
    if (!nquartets || !edgeL) { ctx->err = "bad arguments to snaq_update_bl"; return SNAQ_EINVAL; }
    const SnaqHostTables &H = ctx->H;
    if (ctx->world > 1) { ctx->err = "snaq_update_bl needs an unsharded context"; return SNAQ_EINVAL; }
    if (H.n_hybrids > 0) { ctx->err = "updateBL! was created for a tree, and net here is not a tree, so no branch lengths updated (src/readquartetdata.jl:839-841)"; return SNAQ_EINVAL; }
    CK(cudaSetDevice(ctx->device));
    const int np = H.nparams;
    for (int i = 0; i < H.nt; i++) { nquartets[i] = 0; edgeL[i] = 0.0; }
    if (np == 0 || ctx->nq == 0) return SNAQ_OK;
    unsigned long long *d_bins = nullptr;
    CK(cudaMalloc((void **)&d_bins, (size_t)np * 2 * sizeof(unsigned long long)));
    cudaError_t e = cudaMemsetAsync(d_bins, 0, (size_t)np * 2 * sizeof(unsigned long long), ctx->stream);
    if (e == cudaSuccess && ctx->nA1 > 0) {
        k_update_bl<<<(unsigned)((ctx->nA1 + 255) / 256), 256, 0, ctx->stream>>>(ctx->T, reinterpret_cast<const uint2 *>(ctx->d_prog), ctx->d_qhdr,
                                                                                 ctx->d_perm, ctx->d_obs, (uint32_t)ctx->nA1, d_bins);
        e = cudaGetLastError();
    }
    std::vector<unsigned long long> hb((size_t)np * 2);
    if (e == cudaSuccess) e = cudaMemcpyAsync(hb.data(), d_bins, hb.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_bins);
    if (e != cudaSuccess) { ctx->err = cudaGetErrorString(e); return SNAQ_ECUDA; }
    ctx->counters[5]++;
    ctx->counters[3] += (int64_t)hb.size() * 8;
    for (int i = 0; i < H.nt; i++) {
        const int slot = H.nh + i;
        const unsigned long long n = hb[2 * slot];
        nquartets[i] = (int64_t)n;
        if (n > 0) {
            const double mean = (double)(long long)hb[2 * slot + 1] / UBL_SCALE / (double)n;
            edgeL[i] = -std::log(1.5 * (1.0 - mean));              // :846
        }
    }
    return SNAQ_OK;
}

int snaq_eval_quartets(snaq_ctx *ctx, int32_t nparams, const double *x, double *expcf, double *logpl, double *deltacf) {
    if (!ctx) return SNAQ_EINVAL;
    if (!ctx->has_net) { ctx->err = "snaq_set_network has not been called"; return SNAQ_ENODATA; }
    if (!x || nparams != ctx->H.nparams) { ctx->err = "bad arguments to snaq_eval_quartets (x has wrong length?)"; return SNAQ_EINVAL; }
    int rc = check_x_host(ctx, 1, x);
    if (rc) return rc;
    CK(cudaSetDevice(ctx->device));
    const int64_t nq = ctx->nq;
    rc = ensure(ctx, ctx->d_X, ctx->X_cap, (size_t)std::max(nparams, 1));
    if (rc) return rc;
    rc = ensure(ctx, ctx->d_out, ctx->out_cap, 2);
    if (rc) return rc;
    // the per-quartet arrays live in buffers of the context (a read-back per search step should not pay cudaMalloc)
    double *d_e = nullptr, *d_l = nullptr, *d_d = nullptr;
    const size_t n = (size_t)std::max<int64_t>(nq, 1);
    if (expcf) { rc = ensure(ctx, ctx->d_rb_e, ctx->rb_e_cap, n * 3); if (rc) return rc; d_e = ctx->d_rb_e; }
    if (logpl || (!expcf && !deltacf)) { rc = ensure(ctx, ctx->d_rb_l, ctx->rb_l_cap, n); if (rc) return rc; d_l = ctx->d_rb_l; }
    if (deltacf) { rc = ensure(ctx, ctx->d_rb_d, ctx->rb_d_cap, n); if (rc) return rc; d_d = ctx->d_rb_d; }   // (at least one per-quartet pointer selects the thread-per-quartet kernel)
    CK(cudaMemcpyAsync(ctx->d_X, x, sizeof(double) * nparams, cudaMemcpyHostToDevice, ctx->stream));
    rc = eval_device(ctx, 1, ctx->d_X, ctx->d_out, ctx->stream, d_e, d_l, d_d);
    uint32_t st = 0;
    if (!rc) {
        cudaMemcpyAsync(&st, ctx->d_status + 1, sizeof(st), cudaMemcpyDeviceToHost, ctx->stream);
        if (expcf && nq) cudaMemcpyAsync(expcf, d_e, sizeof(double) * 3 * nq, cudaMemcpyDeviceToHost, ctx->stream);
        if (logpl && nq) cudaMemcpyAsync(logpl, d_l, sizeof(double) * nq, cudaMemcpyDeviceToHost, ctx->stream);
        if (deltacf && nq) cudaMemcpyAsync(deltacf, d_d, sizeof(double) * nq, cudaMemcpyDeviceToHost, ctx->stream);
        cudaError_t e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { ctx->err = cudaGetErrorString(e); rc = SNAQ_ECUDA; }
        ctx->counters[2] += 8 * nparams;
        ctx->counters[3] += (expcf ? 24 : 0) * nq + (logpl ? 8 : 0) * nq + (deltacf ? 8 : 0) * nq;
    }
    if (rc) return rc;
    if (st) { cudaMemsetAsync(ctx->d_status + 1, 0, sizeof(uint32_t), ctx->stream); return status_to_error(ctx, st); }
    return SNAQ_OK;
}

int snaq_sample_quartets(snaq_ctx *ctx, int32_t nparams, const double *x, int32_t n, const double *u, int64_t *idx) {
    if (!ctx) return SNAQ_EINVAL;
    if (!ctx->has_net) { ctx->err = "snaq_set_network has not been called"; return SNAQ_ENODATA; }
    if (!x || nparams != ctx->H.nparams || n < 0 || (n > 0 && (!u || !idx))) { ctx->err = "bad arguments to snaq_sample_quartets"; return SNAQ_EINVAL; }
    if (ctx->world > 1) { ctx->err = "snaq_sample_quartets is not available in quartet-sharded mode"; return SNAQ_EINVAL; }
    for (int k = 0; k < n; k++)
        if (!(u[k] >= 0.0 && u[k] < 1.0)) { ctx->err = "uniform draws must be in [0,1)"; return SNAQ_EINVAL; }
    if (ctx->nq == 0) { ctx->err = "no quartets to sample from"; return SNAQ_ENODATA; }
    int rc = check_x_host(ctx, 1, x);
    if (rc) return rc;
    if (n == 0) return SNAQ_OK;
    CK(cudaSetDevice(ctx->device));
    const int64_t nq = ctx->nq;
    const int nchunks = (int)((nq + SQ_CHUNK - 1) / SQ_CHUNK);
    rc = ensure(ctx, ctx->d_X, ctx->X_cap, (size_t)std::max(nparams, 1));
    if (rc) return rc;
    rc = ensure(ctx, ctx->d_out, ctx->out_cap, 2);
    if (rc) return rc;
    // the weights, the chunk totals and the draws live in buffers of the context: a search proposes thousands of moves
    rc = ensure(ctx, ctx->d_sq_w, ctx->sq_w_cap, (size_t)nq);
    if (rc) return rc;
    rc = ensure(ctx, ctx->d_sq_cum, ctx->sq_cum_cap, (size_t)nchunks);
    if (rc) return rc;
    rc = ensure(ctx, ctx->d_sq_u, ctx->sq_u_cap, (size_t)n);
    if (rc) return rc;
    rc = ensure(ctx, ctx->d_sq_idx, ctx->sq_idx_cap, (size_t)n);
    if (rc) return rc;
    double *d_w = ctx->d_sq_w, *d_cum = ctx->d_sq_cum, *d_u = ctx->d_sq_u;
    long long *d_idx = ctx->d_sq_idx;
    cudaError_t e = cudaSuccess;
    uint32_t st = 0;
    if (e == cudaSuccess) e = cudaMemcpyAsync(ctx->d_X, x, sizeof(double) * nparams, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_u, u, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) {
        rc = eval_device(ctx, 1, ctx->d_X, ctx->d_out, ctx->stream, nullptr, nullptr, d_w);     // deltaCF, the data's quartet order
        if (!rc) {
            k_chunk_sums<<<nchunks, 256, 0, ctx->stream>>>(d_w, nq, d_cum);
            const int in_smem = nchunks <= 24 * 1024 ? 1 : 0;                    // 192 KB of running totals: 100 M quartets
            const size_t smem = in_smem ? (size_t)nchunks * sizeof(double) : 0;
            if (e == cudaSuccess) e = raise_smem_limit(k_sample, smem);
            k_sample<<<1, 256, smem, ctx->stream>>>(d_w, nq, d_cum, nchunks, in_smem, d_u, n, d_idx);
            e = cudaGetLastError();
            if (e == cudaSuccess) e = cudaMemcpyAsync(&st, ctx->d_status + 1, sizeof(st), cudaMemcpyDeviceToHost, ctx->stream);
            if (e == cudaSuccess) e = cudaMemcpyAsync(idx, d_idx, sizeof(long long) * n, cudaMemcpyDeviceToHost, ctx->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        }
    }
    if (rc) return rc;
    if (e != cudaSuccess) { ctx->err = std::string("snaq_sample_quartets: ") + cudaGetErrorString(e); return SNAQ_ECUDA; }
    ctx->counters[2] += 8 * (nparams + n);
    ctx->counters[3] += 8 * n;
    ctx->counters[5] += 2;
    if (st) { cudaMemsetAsync(ctx->d_status + 1, 0, sizeof(uint32_t), ctx->stream); return status_to_error(ctx, st); }
    return SNAQ_OK;
}

int snaq_get_hasedge(snaq_ctx *ctx, uint8_t *hasedge, int32_t *indexht, int32_t *nindexht) {
    if (!ctx) return SNAQ_EINVAL;
    if (!ctx->has_net) { ctx->err = "snaq_set_network has not been called"; return SNAQ_ENODATA; }
    if (!hasedge) { ctx->err = "bad arguments to snaq_get_hasedge"; return SNAQ_EINVAL; }
    const SnaqHostTables &H = ctx->H;
    CK(cudaSetDevice(ctx->device));
    const int64_t nq = ctx->nq;
    const int np = H.nparams;
    if (nq == 0 || np == 0) { if (nindexht) for (int64_t q = 0; q < nq; q++) nindexht[q] = 0; return SNAQ_OK; }
    // One hybrid node at most: what extractQuartet! leaves behind is a structural property of the quartet (snaq_hasedge.h).
    // Nested cycles: it depends on the order in which the other leaves are deleted, so the pruning is simulated
    // (snaq_hesim.h).  SNAQ_HASEDGE_SIM=1 forces the simulation for every network (the tests compare the two).
    const bool sim = H.n_hybrids > 1 || (getenv("SNAQ_HASEDGE_SIM") && atoi(getenv("SNAQ_HASEDGE_SIM")) != 0);
    const int nz = sim ? H.nhz : 0;
    if (sim && (H.sim.n_nodes > SNAQ_SIM_MAX || H.sim.n_edges > SNAQ_SIM_MAX || nz > 64)) {
        ctx->err = "network too large for the hasEdge read-back (more than " + std::to_string(SNAQ_SIM_MAX) + " nodes or edges)";
        return SNAQ_EINVAL;
    }
    uint8_t *d_he = nullptr, *d_bd1 = nullptr, *d_sim = nullptr;
    int16_t *d_z = nullptr;
    CK(cudaMalloc((void **)&d_he, (size_t)nq * np));
    cudaError_t e = cudaMalloc((void **)&d_bd1, (size_t)nq);
    if (e == cudaSuccess && nz > 0) e = cudaMalloc((void **)&d_z, (size_t)nq * nz * sizeof(int16_t));
    if (e == cudaSuccess && sim) e = cudaMalloc((void **)&d_sim, H.sim_blob.size());
    if (e == cudaSuccess && sim) e = cudaMemcpyAsync(d_sim, H.sim_blob.data(), H.sim_blob.size(), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(ctx->d_status, 0, sizeof(uint32_t), ctx->stream);
    if (e == cudaSuccess) {
        if (sim) {
            const int64_t blocks = std::min<int64_t>((nq + 63) / 64, (int64_t)ctx->sm_count * 16);
            k_hasedge_sim<<<(unsigned)blocks, 64, 0, ctx->stream>>>(H.sim_view(d_sim), ctx->d_taxa, nq, d_he, d_bd1, d_z, nz, ctx->d_status);
        } else {
            k_hasedge<<<(unsigned)((nq + 127) / 128), 128, 0, ctx->stream>>>(ctx->T, ctx->d_taxa, nq, d_he, d_bd1, ctx->d_status);
        }
        e = cudaGetLastError();
    }
    std::vector<uint8_t> hb((size_t)nq);
    std::vector<int16_t> hz((size_t)nq * std::max(nz, 1), -1);
    uint32_t st = 0;
    if (e == cudaSuccess) e = cudaMemcpyAsync(hasedge, d_he, (size_t)nq * np, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(hb.data(), d_bd1, (size_t)nq, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess && nz > 0) e = cudaMemcpyAsync(hz.data(), d_z, (size_t)nq * nz * sizeof(int16_t), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(&st, ctx->d_status, sizeof(st), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_he); cudaFree(d_bd1); cudaFree(d_z); cudaFree(d_sim);
    if (e != cudaSuccess) { ctx->err = cudaGetErrorString(e); return SNAQ_ECUDA; }
    if (st) { ctx->err = build_err_text(st); return SNAQ_ETOPOLOGY; }
    ctx->counters[5]++;
    ctx->counters[3] += (int64_t)nq * (np + 1 + 2 * nz);
    if (indexht || nindexht) {
        // parameters!(qnet,net) (src/snaq_optimization.jl:106-180): positions of x the quartet uses, 1-based: gammas, then
        // edge lengths (both in x order = qnet.edge order), then the gammaz slots in qnet.edge order of the edges that
        // carry them; a quartet that keeps a bad diamond I hybrid lists only its two gammaz (:120-146)
        const int nht = H.nh + H.nt;
        for (int64_t q = 0; q < nq; q++) {
            int n = 0;
            const uint8_t *he = hasedge + (size_t)q * np;
            int32_t *ix = indexht ? indexht + (size_t)q * np : nullptr;
            if (!hb[q]) for (int i = 0; i < nht; i++) if (he[i]) { if (ix) ix[n] = i + 1; n++; }
            const int n0 = n;
            for (int i = nht; i < np; i++) if (he[i]) { if (ix) ix[n] = i + 1; n++; }
            if (ix && nz > 0 && !hb[q] && n - n0 > 1) {
                const int16_t *zp = hz.data() + (size_t)q * nz;
                std::stable_sort(ix + n0, ix + n, [&](int32_t a, int32_t b) { return zp[a - 1 - nht] < zp[b - 1 - nht]; });
            }
            if (ix) for (int i = n; i < np; i++) ix[i] = -1;
            if (nindexht) nindexht[q] = n;
        }
    }
    return SNAQ_OK;
}

int snaq_get_quartet_edges(snaq_ctx *ctx, int64_t q, int32_t *edge_numbers, int32_t cap, int32_t *n) {
    if (!ctx) return SNAQ_EINVAL;
    if (!ctx->has_net) { ctx->err = "snaq_set_network has not been called"; return SNAQ_ENODATA; }
    if (!n || cap < 0 || (cap > 0 && !edge_numbers) || q < ctx->q_begin || q >= ctx->q_begin + ctx->nq) {
        ctx->err = "bad arguments to snaq_get_quartet_edges (quartet index outside this context's block?)";
        return SNAQ_EINVAL;
    }
    const SnaqHostTables &H = ctx->H;
    if (H.sim.n_nodes > SNAQ_SIM_MAX || H.sim.n_edges > SNAQ_SIM_MAX) { ctx->err = "network too large for the quartet read-back"; return SNAQ_EINVAL; }
    CK(cudaSetDevice(ctx->device));
    const int capd = std::max(H.sim.n_edges, 1);
    uint8_t *d_sim = nullptr;
    int32_t *d_out = nullptr;
    CK(cudaMalloc((void **)&d_sim, H.sim_blob.size()));
    cudaError_t e = cudaMalloc((void **)&d_out, ((size_t)capd + 1) * sizeof(int32_t));
    std::vector<int32_t> h((size_t)capd + 1, 0);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_sim, H.sim_blob.data(), H.sim_blob.size(), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) {
        k_quartet_edges<<<1, 1, 0, ctx->stream>>>(H.sim_view(d_sim), ctx->d_taxa, q - ctx->q_begin, d_out + 1, capd, d_out);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(h.data(), d_out, h.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_sim); cudaFree(d_out);
    if (e != cudaSuccess) { ctx->err = cudaGetErrorString(e); return SNAQ_ECUDA; }
    if (h[0] < 0) { ctx->err = build_err_text(SNAQ_B_BLOCKS); return SNAQ_ETOPOLOGY; }
    *n = h[0];
    for (int i = 0; i < std::min(h[0], cap); i++) edge_numbers[i] = h[(size_t)i + 1];
    ctx->counters[5]++;
    return SNAQ_OK;
}

int snaq_get_descriptors(snaq_ctx *ctx, int8_t *which, int8_t *ntype, int8_t *types, int32_t max_types, int8_t *formula,
                         uint8_t *hasedge) {
    if (!ctx) return SNAQ_EINVAL;
    if (!ctx->has_net) { ctx->err = "snaq_set_network has not been called"; return SNAQ_ENODATA; }
    if (hasedge) { int rc = snaq_get_hasedge(ctx, hasedge, nullptr, nullptr); if (rc) return rc; }
    CK(cudaSetDevice(ctx->device));
    const int64_t nq = ctx->nq;
    if (nq == 0) return SNAQ_OK;
    const int nh = std::max(ctx->H.n_hybrids, 1);
    int8_t *d_which = nullptr, *d_types = nullptr, *d_formula = nullptr;
    CK(cudaMalloc((void **)&d_which, (size_t)nq));
    CK(cudaMalloc((void **)&d_types, (size_t)nq * nh));
    CK(cudaMalloc((void **)&d_formula, (size_t)nq * 3));
    k_describe<<<(unsigned)((nq + 127) / 128), 128, 0, ctx->stream>>>(ctx->T, ctx->d_taxa, nq, d_which, d_types, ctx->H.n_hybrids, d_formula);
    cudaError_t e = cudaGetLastError();
    std::vector<int8_t> hw((size_t)nq), ht((size_t)nq * nh), hf((size_t)nq * 3);
    if (e == cudaSuccess) e = cudaMemcpyAsync(hw.data(), d_which, hw.size(), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(ht.data(), d_types, ht.size(), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(hf.data(), d_formula, hf.size(), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_which); cudaFree(d_types); cudaFree(d_formula);
    if (e != cudaSuccess) { ctx->err = cudaGetErrorString(e); return SNAQ_ECUDA; }
    ctx->counters[5]++;
    ctx->counters[3] += (int64_t)nq * (4 + nh);
    for (int64_t q = 0; q < nq; q++) {
        if (which) which[q] = hw[q];
        if (formula) for (int i = 0; i < 3; i++) formula[3 * q + i] = hf[3 * q + i];
        int n = 0;
        for (int i = 0; i < ctx->H.n_hybrids; i++) {
            int8_t t = ht[(size_t)q * nh + i];
            if (t > 0) { if (types && n < max_types) types[q * max_types + n] = t; n++; }
        }
        if (types) for (int i = n; i < max_types; i++) types[q * max_types + i] = 0;
        if (ntype) ntype[q] = (int8_t)n;
    }
    return SNAQ_OK;
}

int snaq_measure_fp64_peak(snaq_ctx *ctx, double *tflops) {
    if (!ctx || !tflops) return SNAQ_EINVAL;
    CK(cudaSetDevice(ctx->device));
    int rc = ensure(ctx, ctx->d_out, ctx->out_cap, 2);
    if (rc) return rc;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const int blocks = ctx->sm_count * 8, iters = 1 << 15;
    double best = 0.0;
    for (int rep = 0; rep < 6; rep++) {
        CK(cudaEventRecord(e0, ctx->stream));
        k_dfma_peak<<<blocks, 256, 0, ctx->stream>>>(ctx->d_out, iters, 0.999999, 1e-9);
        CK(cudaEventRecord(e1, ctx->stream));
        CK(cudaEventSynchronize(e1));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        double tf = 2.0 * 8.0 * (double)iters * blocks * 256.0 / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *tflops = best;
    return SNAQ_OK;
}

int snaq_get_trace(snaq_ctx *ctx, uint64_t *out, int32_t max_blocks, int32_t *nblocks) {
    if (!ctx || !nblocks) return SNAQ_EINVAL;
    *nblocks = 0;
    if (!ctx->d_trace) return SNAQ_OK;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    const int n = std::min(ctx->trace_blocks, (int)max_blocks);
    if (n > 0 && out) CK(cudaMemcpy(out, ctx->d_trace, (size_t)n * 8 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    *nblocks = n;
    return SNAQ_OK;
}

int snaq_get_counters(snaq_ctx *ctx, int64_t out[8]) {
    if (!ctx || !out) return SNAQ_EINVAL;
    for (int i = 0; i < 8; i++) out[i] = ctx->counters[i];
    return SNAQ_OK;
}

}  // extern "C"
