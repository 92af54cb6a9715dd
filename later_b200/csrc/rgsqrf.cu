// Recursion driver for RGSQRF and the extern "C" entry points of include/later_b200.h.
//
// The recursion is the reference's (QR/later_rgsqrf.cu:25-60): split the columns in half, factor
// the left half, R12 = Q1^T A2, A2 -= Q1 R12, factor the right half; 128-column base case.  What
// differs is everything underneath:
//   * operands of the two products are fp16 exactly as in the reference (RN cast of Q1, A2 and of
//     R12), but the casts are fused into the producers: the panel kernel and the update kernel
//     write an fp16 shadow of every column they finalise, the gram kernel emits fp16 R12 from its
//     epilogue.  The reference's three s2h passes per node (util/util.cu:24-32) disappear; the
//     caller's input columns are cast once, by the left-spine node that first reads them.
//   * the products are hand-written tcgen05 kernels (tc_gemm.cu, tc_update.cu), not cublasGemmEx.
//   * the base case is one Gram/Cholesky panel (panel.cu; tensor-core kernels of panel_tc.cu for
//     tall panels), not the 26-launch CAQR chain.
//   * the whole launch sequence (893 launches for 16384^2) is captured into a CUDA graph on the
//     second identical call and replayed; the reference serialises ~4 000 launches on the legacy
//     stream.
//   * the host entry point runs the same operations left-looking over column pieces so that the
//     PCIe transfers stream underneath (HostPipe below).
#include "../../include/later_b200.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "context.h"
#include "launch.cuh"
#include "panel.cuh"
#include "tc_gemm.cuh"

namespace lb {

int fail(later_b200_ctx* ctx, int code, const std::string& msg) {
    if (ctx) ctx->error = msg;
    return code;
}
int cuda_fail(later_b200_ctx* ctx, cudaError_t e, const char* where) {
    if (ctx) ctx->error = std::string(where) + ": " + cudaGetErrorString(e);
    (void)cudaGetLastError();
    return (int)e;
}

namespace {

constexpr int NMIN = kPanelWidth;  // base-case width, reference QR/later_rgsqrf.cu:23

inline long round_up(long x, long a) { return (x + a - 1) / a * a; }

// fp32 -> fp16 (RN) of columns [c0, n) of A into the shadow; 8 elements per thread when aligned.
__global__ void cast_shadow_kernel(const float* __restrict__ A, long lda, int m, int c0, int n,
                                   __half* __restrict__ Qh, long ldh, int vec_ok) {
    const int col = c0 + blockIdx.y;
    pdl_trigger();
    pdl_wait();
    if (col >= n) return;
    const float* src = A + (long)col * lda;
    __half* dst = Qh + (long)col * ldh;
    if (vec_ok) {
        for (int i = (blockIdx.x * blockDim.x + threadIdx.x) * 8; i < m;
             i += gridDim.x * blockDim.x * 8) {
            const float4 a = *reinterpret_cast<const float4*>(src + i);
            const float4 b = *reinterpret_cast<const float4*>(src + i + 4);
            __half2 h0 = __floats2half2_rn(a.x, a.y), h1 = __floats2half2_rn(a.z, a.w);
            __half2 h2 = __floats2half2_rn(b.x, b.y), h3 = __floats2half2_rn(b.z, b.w);
            uint4 out;
            out.x = *reinterpret_cast<uint32_t*>(&h0);
            out.y = *reinterpret_cast<uint32_t*>(&h1);
            out.z = *reinterpret_cast<uint32_t*>(&h2);
            out.w = *reinterpret_cast<uint32_t*>(&h3);
            *reinterpret_cast<uint4*>(dst + i) = out;
        }
    } else {
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x)
            dst[i] = __float2half_rn(src[i]);
    }
}

struct Workspace {
    size_t qh_bytes, r12h_bytes, wh_bytes, part_bytes, panel_bytes, stage_bytes, total;
    long ldh;
};

int gram_bn(int h) { return h >= 512 ? 256 : 128; }
// update kernel per level, from a sweep on B200 at m = 16384 (scripts/gpu_upd_variants.py):
//   h <= 2048 : TMA-streamed C, BN = 128 (4-slot C ring)      75 / 183 us at h = 1024 / 2048
//   h  = 4096 : per-thread epilogue, BN = 256 (all smem to the A/B ring)            462 us
//   h >= 8192 : TMA-streamed C, BN = 256                                            1.60 ms
int update_bn(int h) { return h >= 4096 ? 256 : 128; }
bool update_uses_tma(int h) { return h < 4096 || h >= 8192; }

Workspace plan_workspace(int num_sms, int m, int n, bool dist = false) {
    Workspace w{};
    w.ldh = round_up(m, 8);
    w.qh_bytes = (size_t)w.ldh * n * sizeof(__half);
    w.r12h_bytes = n > NMIN ? (size_t)(n / 2) * (n / 2) * sizeof(__half) : 0;
    size_t part = 0;
    for (int h = NMIN; h * 2 <= n; h *= 2) {
        int s = choose_gram_splits(num_sms, h, h, gram_bn(h), m);
        if (s > 1 && tc_gram_cast_supports(h)) s = std::max(s, tc_gram_cast_splits(num_sms, h, m));
        if (s > 1) part = std::max(part, (size_t)s * h * h * sizeof(float));
    }
    for (int h = NMIN; h <= 2 * NMIN && h * 2 <= n; h *= 2)
        if (tc_node_supports(num_sms, m, h)) part = std::max(part, tc_node_part_floats(m, h) * sizeof(float));
    w.part_bytes = part;
    w.panel_bytes = panel_scratch_bytes(m, num_sms);
    w.wh_bytes = (size_t)n * n * sizeof(__half);  // fp16 W of the TSQR back-multiplication
    w.stage_bytes = dist && n > NMIN ? (size_t)(n / 2) * (n / 2) * sizeof(float) : 0;   // all-reduce staging of R12
    w.total = round_up(w.qh_bytes, 256) + round_up(w.r12h_bytes, 256) + round_up(w.wh_bytes, 256) +
              round_up(w.part_bytes, 256) + round_up(w.panel_bytes, 256) + round_up(w.stage_bytes, 256) + 4096;
    return w;
}

// The PCIe legs of later_b200_rgsqrf_host.  The host path runs the factorisation LEFT-LOOKING over
// column pieces of width max(min(n, 256), n/16): piece j is touched only once it has arrived, receives the
// Gram/update of every node of the recursion tree whose right half contains it (widest first - the
// order the recursion applies them in), is then factored by the ordinary recursion, and is final.
// Same operations on the same operands as the recursive order, so the result is bit-identical to
// the device entry point; but everything left of the piece is final when the piece starts and the
// piece is final when it ends, so A streams in, and Q and R stream out, at PCIe speed with the
// factorisation hidden underneath.
// The copies run on two side streams forked from and joined back into the main stream, so the
// same enqueue code works directly and inside a stream capture (where they become memcpy nodes).
struct HostPipe {
    later_b200_ctx* ctx = nullptr;
    float* hA = nullptr; long hlda = 0;
    float* hR = nullptr; long hldr = 0;
    int m = 0, n = 0, chunk = NMIN;
    float* dA = nullptr; long ldda = 0;   // device matrices the factorisation works on
    float* dR = nullptr; long lddr = 0;
    bool copy_out = true;      // false: Q and R stay on the device (later_b200_rgsqrf_stream_in)
    std::vector<int> in_end;   // right edge of each H2D piece
    std::vector<cudaEvent_t> in_ev;
    size_t used = 0;
    cudaError_t err = cudaSuccess;

    void check(cudaError_t e) {
        if (err == cudaSuccess && e != cudaSuccess) err = e;
    }
    cudaEvent_t event() {
        if (used == ctx->events.size()) {
            cudaEvent_t ev = nullptr;
            check(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            if (!ev) return nullptr;
            ctx->events.push_back(ev);
        }
        return ctx->events[used++];
    }
    // (at least 256 columns, so that the nodes of half-width 128 and 256 see their whole right half in one
    // call, as in the recursive order: those are the nodes one fused kernel handles, tc_node)
    static int chunk_for(int n) { return std::max(std::min(n, 2 * NMIN), n / 16); }

    void start() {
        chunk = chunk_for(n);
        cudaEvent_t fork = event();
        if (err != cudaSuccess) return;
        check(cudaEventRecord(fork, ctx->stream));
        check(cudaStreamWaitEvent(ctx->s_in, fork, 0));
        check(cudaStreamWaitEvent(ctx->s_out, fork, 0));
        const size_t col = (size_t)m * sizeof(float);
        for (int c0 = 0; c0 < n && err == cudaSuccess; c0 += chunk) {
            check(cudaMemcpy2DAsync(dA + (size_t)c0 * ldda, (size_t)ldda * sizeof(float),
                                    hA + (size_t)c0 * hlda, (size_t)hlda * sizeof(float), col, chunk,
                                    cudaMemcpyHostToDevice, ctx->s_in));
            cudaEvent_t ev = event();
            if (err != cudaSuccess) return;
            check(cudaEventRecord(ev, ctx->s_in));
            in_end.push_back(c0 + chunk);
            in_ev.push_back(ev);
        }
    }
    // the main stream may read columns [0, c1) after this
    void need(int c1) {
        for (size_t k = 0; k < in_end.size(); ++k)
            if (in_end[k] >= c1) {
                check(cudaStreamWaitEvent(ctx->stream, in_ev[k], 0));
                return;
            }
    }
    // piece [c0, c0 + w) is final: its Q columns and rows [0, c0 + w) of its R columns go back
    void cols_final(int c0, int w) {
        if (!copy_out) return;
        cudaEvent_t ev = event();
        if (err != cudaSuccess) return;
        check(cudaEventRecord(ev, ctx->stream));
        check(cudaStreamWaitEvent(ctx->s_out, ev, 0));
        const size_t col = (size_t)m * sizeof(float);
        check(cudaMemcpy2DAsync(hA + (size_t)c0 * hlda, (size_t)hlda * sizeof(float),
                                dA + (size_t)c0 * ldda, (size_t)ldda * sizeof(float), col, w,
                                cudaMemcpyDeviceToHost, ctx->s_out));
        check(cudaMemcpy2DAsync(hR + (size_t)c0 * hldr, (size_t)hldr * sizeof(float),
                                dR + (size_t)c0 * lddr, (size_t)lddr * sizeof(float),
                                (size_t)(c0 + w) * sizeof(float), w, cudaMemcpyDeviceToHost, ctx->s_out));
    }
    void finish() {
        need(n);   // joins s_in (already waited for by the last piece unless something failed)
        cudaEvent_t ev = event();
        if (err != cudaSuccess) return;
        check(cudaEventRecord(ev, ctx->s_out));
        check(cudaStreamWaitEvent(ctx->stream, ev, 0));
    }
};

// Row-sharded factorisation: the all-reduced R12 block leaves its contiguous staging buffer for R
// (fp32, ld ldr), for the fp16 operand of the update, and the mirror block is cleared.
__global__ void finish_r12_kernel(const float* __restrict__ S, int h, int nb, float* __restrict__ R12, long ldr,
                                  __half* __restrict__ R12h, long ldh, float* __restrict__ Z) {
    const long total = (long)h * nb;
    pdl_trigger();
    pdl_wait();
    for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int i = (int)(idx % h), j = (int)(idx / h);
        const float v = S[idx];
        R12[i + (long)j * ldr] = v;
        R12h[i + (long)j * ldh] = __float2half_rn(v);
        if (Z) Z[i + (long)j * ldr] = 0.f;
    }
}

cudaError_t comm_allreduce(later_b200_ctx* ctx, void* buf, size_t count, int dtype, cudaStream_t stream,
                           const int* only_if = nullptr) {
    // small blocks (everything this factorisation exchanges unless n is huge): one launch of the NVLink
    // peer-memory kernel (peer_comm.cu), ~5 us; NCCL (~33 us per call at 8 GPUs) for the rest
    const size_t bytes = count * (dtype == kNcclDouble ? sizeof(double) : sizeof(float));
    if (ctx->peer.fits(bytes) && ctx->opts.peer_allreduce)
        return ctx->peer.allreduce(buf, count, dtype == kNcclDouble, stream, only_if);
    if (!ctx->comm || !ctx->nccl) return cudaErrorNotReady;
    const ncclResult_t r = ctx->comm_group
        ? ctx->comm_group->allreduce(ctx->nccl, ctx->rank, buf, count, dtype, ctx->comm, stream)
        : ctx->nccl->AllReduce(buf, buf, count, dtype, kNcclSum, ctx->comm, stream);
    if (r != 0) {
        ctx->error = r < 0 ? std::string("another rank of this process gave up")
                           : std::string("ncclAllReduce: ") + ctx->nccl->GetErrorString(r);
        return cudaErrorUnknown;
    }
    return cudaSuccess;
}
cudaError_t comm_allreduce_f64(void* self, double* buf, size_t count, cudaStream_t stream, const int* only_if) {
    return comm_allreduce(static_cast<later_b200_ctx*>(self), buf, count, kNcclDouble, stream, only_if);
}

struct Recursion {
    later_b200_ctx* ctx;
    later_b200_ctx::Plan* p;
    CUtensorMap q128, q256, q64;
    cudaError_t err = cudaSuccess;
    long launches = 0;
    int colmax_col = -1;    // first column of the panel whose column maxima the last update left behind

    void check(cudaError_t e) {
        if (err == cudaSuccess && e != cudaSuccess) err = e;
    }

    void qr(int c0, int w) {
        if (err != cudaSuccess) return;
        cudaStream_t st = ctx->stream;
        if (w <= NMIN) {
            const bool ready = colmax_col == c0;
            colmax_col = -1;
            const PanelComm comm{ctx, comm_allreduce_f64};
            check(panel_qr128(st, ctx->num_sms, p->m, p->A + (long)c0 * p->lda, p->lda,
                              p->R + c0 + (long)c0 * p->ldr, p->ldr, p->Qh + (long)c0 * p->ldh,
                              p->ldh, p->panel_scratch, true, ctx->opts.panel, ctx->d_info, c0, ready,
                              p->dist ? &comm : nullptr));
            launches += panel_launch_count(p->m, ctx->num_sms, p->A + (long)c0 * p->lda, p->lda, true,
                                           ctx->opts.panel) - (ready ? 1 : 0);
        } else {
            qr(c0, w / 2);
            node_tail(c0, w);
        }
    }

    void cast(int c0, int c1) {
        const int vec_ok = (p->lda % 4 == 0) && ((reinterpret_cast<uintptr_t>(p->A) & 15) == 0);
        dim3 grid((unsigned)std::min<long>((p->m / 8 + 255) / 256, 64), (unsigned)(c1 - c0));
        check(launch_pdl(cast_shadow_kernel, grid, dim3(256), 0, ctx->stream, (const float*)p->A,
                         (long)p->lda, p->m, c0, c1, p->Qh, p->ldh, vec_ok));
        launches += 1;
    }

    // R12 = Q1^T A2 and A2 -= Q1 R12 for Q1 = columns [c0, c0 + h), A2 = columns [cb, cb + nb).
    // The recursion calls it with A2 = the whole right half of node (c0, 2h); the left-looking
    // host schedule with one piece of it.  The split-K factor is always the one of the whole node,
    // so that both orders add the same partial sums in the same order.
    // b_is_input: A2 is still the caller's fp32 input - no kernel has written its fp16 shadow yet.
    // Tall (bandwidth-bound) products then round it to fp16 inside the Gram kernel's load path
    // (tc_gram_cast.cu); otherwise the shadow is made first.
    // panel_next: columns cb .. cb + 127 are factored next (their fp16 shadow would never be read).
    void gram_update(int c0, int h, int cb, int nb, bool zero_mirror, bool b_is_input, bool panel_next) {
        if (err != cudaSuccess) return;
        cudaStream_t st = ctx->stream;
        // A small node whose right half is factored next, on a matrix short enough for one CTA per 128-row
        // tile: Gram product, reduce and update in one launch (tc_node, tc_update.cu).
        if (h <= ctx->opts.node_fuse && nb == h && panel_next && !p->dist && tc_node_supports(ctx->num_sms, p->m, h) &&
            p->lda % 4 == 0 && (reinterpret_cast<uintptr_t>(p->A) & 15) == 0) {
            if (b_is_input) cast(cb, cb + nb);
            check(tc_node(st, ctx->num_sms, q128, p->m, h, c0, cb, p->A, p->n, p->lda, p->Qh, p->ldh,
                          p->R + c0 + (long)cb * p->ldr, p->ldr,
                          zero_mirror ? p->R + cb + (long)c0 * p->ldr : nullptr, p->R12h, p->part,
                          ctx->d_info + kInfoWords, ctx->opts.node_coop));
            launches += 1;
            colmax_col = -1;
            return;
        }
        const int bn = nb % 256 == 0 ? gram_bn(h) : 128;
        // Tall nodes on the left spine take the split-K factor of the cast-fused kernel (one that fills the SMs
        // with its Mc x 128 strips) whether or not that kernel is enabled: the factor fixes the summation order.
        const int generic = choose_gram_splits(ctx->num_sms, h, h, gram_bn(h), p->m);
        const bool cast_node = b_is_input && generic >= 2 && tc_gram_cast_supports(h) && p->m >= kTcApplyMinRows &&
                               p->lda % 4 == 0 && (reinterpret_cast<uintptr_t>(p->A) & 15) == 0;
        const int splits = cast_node ? tc_gram_cast_splits(ctx->num_sms, h, p->m) : generic;
        float* R12 = p->R + c0 + (long)cb * p->ldr;
        // (the mirror block R21, which the algorithm never produces, is written as zero on the way)
        float* Z = zero_mirror ? p->R + cb + (long)c0 * p->ldr : nullptr;
        // Row-sharded: this rank's product is only its share of R12 = sum over ranks of Q1_p^T A2_p.  It goes
        // to a contiguous staging block, is summed over the ranks (NCCL all-reduce, fp32) and only then
        // leaves for R and for the fp16 operand of the update - every rank continues with the same R12.
        float* C = p->dist ? p->stage : R12;
        const long ldc = p->dist ? h : p->ldr;
        __half* Ch = p->dist ? nullptr : p->R12h;
        float* Zk = p->dist ? nullptr : Z;
        const bool fused = cast_node && ctx->opts.gram_cast;
        if (fused) {
            check(tc_gram_cast(st, ctx->num_sms, q128, p->m, c0, h, p->A + (long)cb * p->lda, p->lda, nb,
                               C, ldc, Ch, h, p->part, splits, Zk));
        } else {
            if (b_is_input) cast(cb, cb + nb);
            check(tc_gram(st, ctx->num_sms, q128, bn == 256 ? q256 : q128, bn, 0, p->m, c0, h, cb, nb, C,
                          ldc, Ch, h, p->part, splits, Zk));
        }
        launches += splits > 1 ? 2 : 1;
        if (p->dist) {
            check(comm_allreduce(ctx, p->stage, (size_t)h * nb, kNcclFloat, st));
            const int blocks = (int)std::min<long>(((long)h * nb + 255) / 256, 148L * 8);
            check(launch_pdl(finish_r12_kernel, dim3(blocks), dim3(256), 0, st, (const float*)p->stage, h, nb, R12,
                             (long)p->ldr, p->R12h, (long)h, Z));
            launches += 1;
        }
        CUtensorMap r12map;
        HalfMatrix rm{p->R12h, h, nb, h};
        const int ubn = nb % 256 == 0 ? update_bn(h) : 128;
        check(make_tensor_map_f16(&r12map, rm, 64, ubn));
        colmax_col = -1;
        if (update_uses_tma(h) && p->lda % 4 == 0 && (reinterpret_cast<uintptr_t>(p->A) & 15) == 0) {
            // columns cb .. cb + 127 are the next panel: if it will use the integer Gram kernel, this
            // update leaves their maxima behind and saves that kernel its first pass over the panel
            // (~2 % more epilogue work against a separate launch that costs 15 us even when the panel
            // still sits in L2: 5 % of the step on a 131072-row shard)
            const bool want = panel_uses_i8_gram(p->m, ctx->num_sms, p->A + (long)cb * p->lda, p->lda, true,
                                                 ctx->opts.panel);
            check(tc_update_tma(st, ctx->num_sms, q64, r12map, ubn, 0, p->m, c0, h, 0, nb, p->A, p->m,
                                p->n, p->lda, cb, p->Qh, p->ldh, true,
                                want ? panel_colmax_scratch(p->panel_scratch, p->m, ctx->num_sms) : nullptr,
                                kColmaxParts, panel_next ? NMIN : 0));
            if (want) colmax_col = cb;
        } else {   // (TMA also needs 16-byte aligned column strides)
            check(tc_update(st, ctx->num_sms, q64, r12map, ubn, 0, p->m, c0, h, 0, nb,
                            p->A + (long)cb * p->lda, p->lda, p->Qh + (long)cb * p->ldh, p->ldh, true));
        }
        launches += 1;
    }

    // The host entry point's schedule (see HostPipe).
    void left_looking(HostPipe* pipe) {
        const int P = pipe->chunk, pieces = p->n / P;
        for (int j = 0; j < pieces && err == cudaSuccess; ++j) {
            const int cj = j * P;
            pipe->need(cj + P);
            // (piece 0 is cast by the left spine of its own recursion; a later piece is the caller's
            // input until the first - widest - of these updates has run)
            bool fresh = j > 0;
            int last = 0;                           // the narrowest node that updates this piece
            for (int s = pieces; s >= 2; s /= 2)
                if (j - j / s * s >= s / 2) last = s;
            for (int s = pieces; s >= 2; s /= 2) {
                const int a = j / s * s;            // node (a, s) in pieces; j in its right half?
                if (j - a >= s / 2) {
                    gram_update(a * P, s / 2 * P, cj, P, false, fresh, s == last);
                    fresh = false;
                }
            }
            qr(cj, P);
            // R stays on the device: its blocks below this piece's diagonal block, which no Gram
            // epilogue of this schedule covers, are cleared here (the host copy-out skips them)
            if (!pipe->copy_out && cj + P < p->n)
                check(cudaMemset2DAsync(p->R + (cj + P) + (long)cj * p->ldr, (size_t)p->ldr * sizeof(float),
                                        0, (size_t)(p->n - cj - P) * sizeof(float), P, ctx->stream));
            pipe->cols_final(cj, P);
        }
    }

    // Everything of node (c0, w) after its left recursion: R12 = Q1^T A2, A2 -= Q1 R12, right half.
    void node_tail(int c0, int w) {
        if (err != cudaSuccess) return;
        const int h = w / 2;
        // Left spine (c0 == 0): A2 is still the caller's input; everywhere else the update that last
        // wrote A2 has refreshed its fp16 shadow.
        gram_update(c0, h, c0 + h, h, true, c0 == 0, true);
        qr(c0 + h, h);
    }
};

int validate(later_b200_ctx* ctx, int m, int n, const void* A, int lda, const void* R, int ldr) {
    if (!ctx) return LATER_B200_EINVAL;
    if (!A || !R) return fail(ctx, LATER_B200_EINVAL, "null matrix pointer");
    // the halving recursion reaches the 128-column base case only from n = 128 * 2^k (the reference's
    // panel rejects anything else, QR/panel.cu:12-15)
    if (n < NMIN || n % NMIN != 0 || ((n / NMIN) & (n / NMIN - 1)) != 0)
        return fail(ctx, LATER_B200_EINVAL, "n must be 128 * 2^k");
    if (m < n) return fail(ctx, LATER_B200_EINVAL, "m must be >= n");
    if (m % 8 != 0) return fail(ctx, LATER_B200_EINVAL, "m must be a multiple of 8");
    if (lda < m || ldr < n) return fail(ctx, LATER_B200_EINVAL, "leading dimension too small");
    if ((long)m * n > (long)1 << 34) return fail(ctx, LATER_B200_EINVAL, "matrix too large");
    if (n > 65536) return fail(ctx, LATER_B200_EINVAL, "n must be <= 65536");   // (grid.y of the cast kernel)
    return 0;
}

// Reserves and carves the workspace for (m, n); fills ctx->plan.
int prepare_plan(later_b200_ctx* ctx, int m, int n, float* A, int lda, float* R, int ldr, bool dist = false) {
    const Workspace w = plan_workspace(ctx->num_sms, m, n, dist);
    cudaError_t e = ctx->arena.reserve(w.total);
    if (e != cudaSuccess) {
        cuda_fail(ctx, e, "workspace reserve");
        return LATER_B200_ENOMEM;
    }
    ctx->arena.reset();
    auto& p = ctx->plan;
    p.m = m; p.n = n; p.A = A; p.lda = lda; p.R = R; p.ldr = ldr;
    p.hA = nullptr; p.hlda = 0; p.hR = nullptr; p.hldr = 0;
    p.ldh = w.ldh;
    p.Qh = static_cast<__half*>(ctx->arena.alloc(w.qh_bytes));
    p.R12h = w.r12h_bytes ? static_cast<__half*>(ctx->arena.alloc(w.r12h_bytes)) : nullptr;
    p.Wh = static_cast<__half*>(ctx->arena.alloc(w.wh_bytes));
    p.part = w.part_bytes ? static_cast<float*>(ctx->arena.alloc(w.part_bytes)) : nullptr;
    p.part_floats = w.part_bytes / sizeof(float);
    p.panel_scratch = ctx->arena.alloc(w.panel_bytes);
    p.dist = dist;
    p.stage = w.stage_bytes ? static_cast<float*>(ctx->arena.alloc(w.stage_bytes)) : nullptr;
    p.arena_gen = ctx->arena.generation();
    p.valid = p.Qh && p.Wh && p.panel_scratch && (!w.r12h_bytes || p.R12h) && (!w.part_bytes || p.part) &&
              (!w.stage_bytes || p.stage);
    if (!p.valid) return fail(ctx, LATER_B200_ENOMEM, "workspace carve failed");
    return 0;
}

// Enqueues one factorisation on ctx->stream (directly, or into an ongoing capture): the plain
// launch sequence for a device matrix, or the same with the PCIe copies of the host entry point
// forked around it.
enum Stage : int { STAGE_ALL = 0, STAGE_HOST = 1, STAGE_DIST = 2 };

int enqueue_stage(later_b200_ctx* ctx, int stage, long* launches) {
    auto& p = ctx->plan;
    Recursion rec{};
    rec.ctx = ctx;
    rec.p = &p;
    HalfMatrix qm{p.Qh, p.m, p.n, p.ldh};
    cudaError_t e;
    if ((e = make_tensor_map_f16(&rec.q128, qm, 64, 128)) != cudaSuccess ||
        (e = make_tensor_map_f16(&rec.q256, qm, 64, 256)) != cudaSuccess ||
        (e = make_tensor_map_f16(&rec.q64, qm, 64, 64)) != cudaSuccess)
        return cuda_fail(ctx, e, "tensor map encode");
    if ((e = cudaMemsetAsync(ctx->d_info, 0, kInfoWords * sizeof(int), ctx->stream)) != cudaSuccess)
        return cuda_fail(ctx, e, "info clear");
    HostPipe pipe;
    if (stage == STAGE_HOST) {
        pipe.ctx = ctx;
        pipe.hA = p.hA; pipe.hlda = p.hlda; pipe.hR = p.hR; pipe.hldr = p.hldr;
        pipe.m = p.m; pipe.n = p.n;
        pipe.dA = p.A; pipe.ldda = p.lda; pipe.dR = p.R; pipe.lddr = p.ldr;
        pipe.copy_out = p.hR != nullptr;
        pipe.start();
        rec.left_looking(&pipe);
    } else {
        rec.qr(0, p.n);
    }
    if (stage == STAGE_HOST) {
        pipe.finish();   // always rejoin the side streams, also after an error (capture must close)
        if (pipe.err != cudaSuccess) return cuda_fail(ctx, pipe.err, "host pipeline");
    }
    if (rec.err != cudaSuccess) return cuda_fail(ctx, rec.err, "rgsqrf enqueue");
    // the status words travel to page-locked host memory behind the last kernel (later_b200_last_info)
    if ((e = cudaMemcpyAsync(ctx->h_info, ctx->d_info, kInfoWords * sizeof(int), cudaMemcpyDeviceToHost,
                             ctx->stream)) != cudaSuccess)
        return cuda_fail(ctx, e, "info read-back");
    e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(ctx, e, "rgsqrf launch");
    *launches = rec.launches;
    return 0;
}

bool same_plan(const later_b200_ctx::Plan& a, const later_b200_ctx::Plan& b) {
    return a.valid && b.valid && a.m == b.m && a.n == b.n && a.A == b.A && a.lda == b.lda &&
           a.R == b.R && a.ldr == b.ldr && a.Qh == b.Qh && a.arena_gen == b.arena_gen && a.dist == b.dist &&
           a.hA == b.hA && a.hlda == b.hlda && a.hR == b.hR && a.hldr == b.hldr;
}

// Runs one stage on ctx->stream: replays the cached graph of that stage when a cached plan (shape
// and every pointer) matches, otherwise launches directly (first sight of a plan) or captures and
// instantiates it (second sight).  kGraphSlots plans are kept per stage, least recently used evicted.
int run_stage(later_b200_ctx* ctx, int stage) {
    cudaError_t e;
    auto direct = [&]() {
        long l = 0;
        int rc = enqueue_stage(ctx, stage, &l);
        ctx->launches = l;
        return rc;
    };
    if (!ctx->use_graph) return direct();
    later_b200_ctx::GraphSlot* slot = nullptr;
    for (auto& g : ctx->graphs[stage])
        if (g.seen && same_plan(g.plan, ctx->plan)) { slot = &g; break; }
    if (!slot) {
        // A plan seen for the first time is launched directly: capturing and instantiating a graph of
        // ~900 nodes costs tens of milliseconds, which only pays off from the second identical call on
        // (the reference's driver, test/test_qr.cu, calls the factorisation exactly once per process).
        int rc = direct();
        later_b200_ctx::GraphSlot* victim = &ctx->graphs[stage][0];
        for (auto& g : ctx->graphs[stage]) {
            if (!g.seen) { victim = &g; break; }
            if (g.tick < victim->tick) victim = &g;
        }
        if (victim->exec) cudaGraphExecDestroy(victim->exec);
        victim->exec = nullptr;
        victim->plan = ctx->plan;
        victim->seen = rc == 0;
        victim->launches = ctx->launches;
        victim->tick = ++ctx->graph_tick;
        return rc;
    }
    slot->tick = ++ctx->graph_tick;
    if (slot->exec) {
        e = cudaGraphLaunch(slot->exec, ctx->stream);
        if (e != cudaSuccess) return cuda_fail(ctx, e, "cudaGraphLaunch");
        ctx->launches = slot->launches;
        ++ctx->graph_replays;
        return 0;
    }
    // Capture on a private stream so the legacy default stream can be the context's stream.
    cudaStream_t cap = nullptr;
    if ((e = cudaStreamCreateWithFlags(&cap, cudaStreamNonBlocking)) != cudaSuccess)
        return cuda_fail(ctx, e, "capture stream");
    cudaStream_t user = ctx->stream;
    ctx->stream = cap;
    long launches = 0;
    e = cudaStreamBeginCapture(cap, cudaStreamCaptureModeThreadLocal);
    if (e != cudaSuccess) {
        ctx->stream = user;
        cudaStreamDestroy(cap);
        return cuda_fail(ctx, e, "begin capture");
    }
    int rc = enqueue_stage(ctx, stage, &launches);
    cudaGraph_t graph = nullptr;
    e = cudaStreamEndCapture(cap, &graph);
    ctx->stream = user;
    cudaStreamDestroy(cap);
    if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (e != cudaSuccess) return cuda_fail(ctx, e, "end capture");
    e = cudaGraphInstantiate(&slot->exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) { slot->exec = nullptr; return cuda_fail(ctx, e, "graph instantiate"); }
    slot->launches = launches;
    ctx->launches = launches;
    ++ctx->graph_captures;
    e = cudaGraphLaunch(slot->exec, ctx->stream);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "cudaGraphLaunch");
    return 0;
}

int rgsqrf_prepare(later_b200_ctx* ctx, int m, int n, float* A, int lda, float* R, int ldr, bool dist = false) {
    int rc = validate(ctx, m, n, A, lda, R, ldr);
    if (rc) return rc;
    return prepare_plan(ctx, m, n, A, lda, R, ldr, dist);
}

int rgsqrf_device(later_b200_ctx* ctx, int m, int n, float* A, int lda, float* R, int ldr) {
    if (!ctx) return LATER_B200_EINVAL;
    DeviceGuard guard(ctx->device);
    if (guard.error() != cudaSuccess) return cuda_fail(ctx, guard.error(), "cudaSetDevice");
    int rc = rgsqrf_prepare(ctx, m, n, A, lda, R, ldr);
    if (rc) return rc;
    return run_stage(ctx, STAGE_ALL);
}

void read_options(Options& o) {
    auto geti = [](const char* name, int dflt) { const char* v = getenv(name); return v ? atoi(v) : dflt; };
    o.panel.apply_tc = geti("LB_APPLY_TC", -1);
    o.panel.gram_i8 = geti("LB_GRAM_I8", 1) != 0;
    o.panel.gram_i8_min_rows = geti("LB_GRAM_I8_MIN_ROWS", kI8GramMinRows);
    if (const char* v = getenv("LB_I8_FALLBACK_TAU")) o.panel.i8_fallback_tau = atof(v);
    o.gram_cast = geti("LB_GRAM_CAST", 1) != 0;
    o.update_variant = geti("LB_UPDATE_VARIANT", 0);
    o.ormqr_kchunk = std::max(64, geti("LB_ORMQR_KCHUNK", 2048) / 64 * 64);
    o.gram_2cta = geti("LB_GRAM_2CTA", 1) != 0;
    o.peer_allreduce = geti("LB_PEER_ALLREDUCE", 1) != 0;
    o.node_fuse = geti("LB_NODE_FUSE", 128);
    o.node_coop = geti("LB_NODE_COOP", 1) != 0;
}

std::string rank_message(const int* info) {
    char buf[200];
    if (info[INFO_BAD_COLUMN])
        snprintf(buf, sizeof buf, "numerically rank-deficient input: non-positive Cholesky pivot at column %d "
                 "(clamped; Q and R are not reliable from that column on)", info[INFO_BAD_COLUMN] - 1);
    else
        snprintf(buf, sizeof buf, "non-finite or out-of-range entries in the input");
    return buf;
}

__global__ void cast_matrix_kernel(const float* __restrict__ S, long lds, int rows, int cols,
                                   __half* __restrict__ D, long ldd) {
    const long total = (long)rows * cols;
    for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total;
         idx += (long)gridDim.x * blockDim.x) {
        const int i = (int)(idx % rows), j = (int)(idx / rows);
        D[i + (long)j * ldd] = __float2half_rn(S[i + (long)j * lds]);
    }
}

}  // namespace
}  // namespace lb

using namespace lb;

extern "C" {

int later_b200_create(later_b200_ctx** out, int device, void* stream) {
    if (!out) return LATER_B200_EINVAL;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) {
        (void)cudaGetLastError();
        return LATER_B200_ENODEV;
    }
    cudaDeviceProp prop{};
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return LATER_B200_ENODEV;
    if (prop.major != 10) return LATER_B200_ENODEV;  // tcgen05 kernels: sm_100a only, no fallback
    DeviceGuard guard(device);
    if (guard.error() != cudaSuccess) return LATER_B200_ENODEV;
    later_b200_ctx* ctx = new (std::nothrow) later_b200_ctx();
    if (!ctx) return LATER_B200_ENOMEM;
    ctx->device = device;
    ctx->num_sms = prop.multiProcessorCount;
    ctx->stream = static_cast<cudaStream_t>(stream);
    ctx->arena.bind(ctx->stream);
    read_options(ctx->opts);
    cudaError_t e = cudaMalloc(&ctx->d_info, (kInfoWords + kSyncWords) * sizeof(int));
    if (e == cudaSuccess) e = cudaMallocHost(&ctx->h_info, kInfoWords * sizeof(int));
    if (e == cudaSuccess) {
        memset(ctx->h_info, 0, kInfoWords * sizeof(int));
        e = cudaMemset(ctx->d_info, 0, (kInfoWords + kSyncWords) * sizeof(int));
    }
    if (e == cudaSuccess) e = tc_gemm_init();
    if (e == cudaSuccess) e = tc_gram_cast_init();
    if (e == cudaSuccess) e = tc_update_init();
    if (e == cudaSuccess) e = tc_node_init();
    if (e == cudaSuccess) e = panel_init();
    if (e != cudaSuccess) {
        int rc = cuda_fail(ctx, e, "context setup");
        if (ctx->d_info) cudaFree(ctx->d_info);
        if (ctx->h_info) cudaFreeHost(ctx->h_info);
        delete ctx;
        return rc;
    }
    *out = ctx;
    return 0;
}

int later_b200_destroy(later_b200_ctx* ctx) {
    if (!ctx) return LATER_B200_EINVAL;
    DeviceGuard guard(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (auto& stage : ctx->graphs)
        for (auto& g : stage)
            if (g.exec) cudaGraphExecDestroy(g.exec);
    if (ctx->s_in) cudaStreamDestroy(ctx->s_in);
    if (ctx->s_out) cudaStreamDestroy(ctx->s_out);
    for (auto& ev : ctx->events)
        if (ev) cudaEventDestroy(ev);
    if (ctx->dA) cudaFree(ctx->dA);
    if (ctx->dR) cudaFree(ctx->dR);
    if (ctx->aux) cudaFree(ctx->aux);
    if (ctx->comm && ctx->nccl) ctx->nccl->CommDestroy(ctx->comm);
    ctx->peer.release();
    if (ctx->d_info) cudaFree(ctx->d_info);
    if (ctx->h_info) cudaFreeHost(ctx->h_info);
    ctx->arena.release();
    cudaStreamSynchronize(ctx->stream);
    delete ctx;
    return 0;
}

const char* later_b200_last_error(const later_b200_ctx* ctx) {
    return ctx ? ctx->error.c_str() : "null context";
}

int later_b200_set_graph(later_b200_ctx* ctx, int enable) {
    if (!ctx) return LATER_B200_EINVAL;
    ctx->use_graph = enable != 0;
    return 0;
}

size_t later_b200_workspace_bytes(const later_b200_ctx* ctx, int m, int n) {
    if (!ctx || n <= 0 || m < n) return 0;
    return plan_workspace(ctx->num_sms, m, n).total;
}

long later_b200_last_launch_count(const later_b200_ctx* ctx) { return ctx ? ctx->launches : 0; }

int later_b200_last_info(later_b200_ctx* ctx, int* info) {
    if (!ctx || !info) return LATER_B200_EINVAL;
    DeviceGuard guard(ctx->device);
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "sync");
    info[0] = ctx->h_info[INFO_BAD_COLUMN];
    info[1] = ctx->h_info[INFO_FLAGS];
    info[2] = ctx->h_info[INFO_FALLBACKS];
    info[3] = ctx->h_info[INFO_COND_LOG2];
    return (info[0] != 0 || (info[1] & 1)) ? LATER_B200_ERANK : 0;
}

int later_b200_graph_stats(const later_b200_ctx* ctx, long* replays, long* captures) {
    if (!ctx) return LATER_B200_EINVAL;
    if (replays) *replays = ctx->graph_replays;
    if (captures) *captures = ctx->graph_captures;
    return 0;
}

int later_b200_rgsqrf(later_b200_ctx* ctx, int m, int n, float* A, int lda, float* R, int ldr) {
    return rgsqrf_device(ctx, m, n, A, lda, R, ldr);
}

// ---- row-sharded factorisation: the same recursion on this rank's row block, with the Gram matrix of
// every panel and every R12 block summed over the ranks (NCCL all-reduce), so that all ranks factor the
// GLOBAL matrix: same algorithm, same accuracy as one GPU, 1/P of the work each, no redundant stack
// factorisation and no back-multiplication.
int later_b200_comm_unique_id(void* id128) {
    if (!id128) return LATER_B200_EINVAL;
    std::string err;
    Nccl* nccl = Nccl::get(&err);
    if (!nccl) { fprintf(stderr, "later_b200: %s\n", err.c_str()); return LATER_B200_ESTATE; }
    NcclUniqueId id;
    if (nccl->GetUniqueId(&id) != 0) return LATER_B200_ESTATE;
    memcpy(id128, id.internal, kNcclUniqueIdBytes);
    return 0;
}

int later_b200_comm_init(later_b200_ctx* ctx, int nranks, int rank, const void* id128) {
    if (!ctx || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return LATER_B200_EINVAL;
    if (ctx->comm) return fail(ctx, LATER_B200_ESTATE, "context already has a communicator");
    std::string err;
    ctx->nccl = Nccl::get(&err);
    if (!ctx->nccl) return fail(ctx, LATER_B200_ESTATE, err);
    DeviceGuard guard(ctx->device);
    NcclUniqueId id;
    memcpy(id.internal, id128, kNcclUniqueIdBytes);
    const ncclResult_t r = ctx->nccl->CommInitRank(&ctx->comm, nranks, id, rank);
    if (r != 0) {
        ctx->comm = nullptr;
        return fail(ctx, LATER_B200_ESTATE, std::string("ncclCommInitRank: ") + ctx->nccl->GetErrorString(r));
    }
    ctx->nranks = nranks;
    ctx->rank = rank;
    return 0;
}

int later_b200_comm_init_all(later_b200_ctx* const* ctxs, int nranks) {
    if (!ctxs || nranks < 1) return LATER_B200_EINVAL;
    std::string err;
    Nccl* nccl = Nccl::get(&err);
    std::vector<int> devices(nranks);
    for (int p = 0; p < nranks; ++p) {
        if (!ctxs[p]) return LATER_B200_EINVAL;
        if (!nccl) return fail(ctxs[p], LATER_B200_ESTATE, err);
        if (ctxs[p]->comm) return fail(ctxs[p], LATER_B200_ESTATE, "context already has a communicator");
        devices[p] = ctxs[p]->device;
    }
    std::vector<ncclComm_t> comms(nranks, nullptr);
    const ncclResult_t r = nccl->CommInitAll(comms.data(), nranks, devices.data());
    if (r != 0) return fail(ctxs[0], LATER_B200_ESTATE, std::string("ncclCommInitAll: ") + nccl->GetErrorString(r));
    if (nranks > CommGroup::kMax) return fail(ctxs[0], LATER_B200_EINVAL, "too many ranks in one process");
    auto group = std::make_shared<CommGroup>();
    group->P = nranks;
    for (int p = 0; p < nranks; ++p) {
        ctxs[p]->nccl = nccl;
        ctxs[p]->comm = comms[p];
        ctxs[p]->nranks = nranks;
        ctxs[p]->rank = p;
        ctxs[p]->comm_group = group;
    }
    return 0;
}

int later_b200_rgsqrf_dist(later_b200_ctx* ctx, int m_local, int n, float* A, int lda, float* R, int ldr) {
    if (!ctx) return LATER_B200_EINVAL;
    if (!ctx->comm && !ctx->peer.ready)
        return fail(ctx, LATER_B200_ESTATE, "later_b200_rgsqrf_dist needs later_b200_comm_init first");
    DeviceGuard guard(ctx->device);
    if (guard.error() != cudaSuccess) return cuda_fail(ctx, guard.error(), "cudaSetDevice");
    int rc = rgsqrf_prepare(ctx, m_local, n, A, lda, R, ldr, true);
    if (rc == 0) {
        // (ranks that are threads of this process meet inside every all-reduce: plain stream launches
        // then - capturing NCCL operations of several devices into separate graphs concurrently is not
        // something to rely on; the one-process-per-GPU form replays graphs)
        const bool graph = ctx->use_graph;
        if (ctx->comm_group) ctx->use_graph = false;
        rc = run_stage(ctx, STAGE_DIST);
        ctx->use_graph = graph;
    }
    if (rc && ctx->comm_group) ctx->comm_group->abort();     // do not leave the other threads waiting
    return rc;
}

// Gram-Schmidt twice ("CGS2-style" re-orthogonalisation): A = Q1 R1, Q1 = Q2 R2,
// R = R2 R1.  The second pass sees an almost orthonormal matrix, so its orthogonality no longer
// depends on cond(A): for the price of a second factorisation and one triangular product (fp32-faithful
// split-precision tcgen05 GEMMs) the result is orthogonal to the fp16 level of a well-conditioned input.
int later_b200_rgsqrf_reorth(later_b200_ctx* ctx, int m, int n, float* A, int lda, float* R, int ldr) {
    int rc = validate(ctx, m, n, A, lda, R, ldr);
    if (rc) return rc;
    DeviceGuard guard(ctx->device);
    cudaError_t e = guard.error();
    if (e != cudaSuccess) return cuda_fail(ctx, e, "cudaSetDevice");
    const size_t nn = (size_t)n * n;
    const size_t need = 2 * nn * sizeof(float) + split_gemm_scratch_bytes(n, n, n) + 256;
    if (ctx->aux_bytes < need) {
        if (ctx->aux) cudaFree(ctx->aux);
        ctx->aux = nullptr; ctx->aux_bytes = 0;
        if ((e = cudaMalloc(&ctx->aux, need)) != cudaSuccess) return cuda_fail(ctx, e, "cudaMalloc scratch");
        ctx->aux_bytes = need;
    }
    float* R1 = static_cast<float*>(ctx->aux);
    float* R2 = R1 + nn;
    void* scratch = R2 + nn;
    if ((rc = rgsqrf_device(ctx, m, n, A, lda, R1, n)) != 0) return rc;
    long launches = ctx->launches;
    if ((rc = rgsqrf_device(ctx, m, n, A, lda, R2, n)) != 0) return rc;
    launches += ctx->launches;
    rc = split_gemm_nn(ctx, n, n, n, R2, n, R1, n, R, ldr, scratch, &launches);
    ctx->launches = launches;
    return rc;
}

// Out-of-core front end (reference QR/later_oc_qr.cu:29-121): the matrix lives in host
// memory and may be larger than the device.  Column blocks of width B stream through a device window of
// three blocks - two slots for finished Q blocks arriving from the host, one for the block being worked on:
//   for block j:  A_j -> device;  for i < j:  Q_i -> device (next slot, while the previous product runs),
//                 R_ij = Q_i^T A_j,  A_j -= Q_i R_ij   (the tcgen05 trailing-update pair);
//                 A_j = Q_j R_jj (the in-core recursion);  Q_j, R_ij, R_jj -> host.
// Block Gram-Schmidt with sequential projections, as the reference's recursion does between its
// 8192-column panels (QR/later_oc_qr.cu:70-88), on the same kernels as the in-core path.
int later_b200_oc_qr(later_b200_ctx* ctx, int m, int n, float* hA, int lda, float* hR, int ldr, int block_cols) {
    if (!ctx) return LATER_B200_EINVAL;
    if (!hA || !hR) return fail(ctx, LATER_B200_EINVAL, "null matrix pointer");
    const int B = block_cols;
    if (B < NMIN || B % NMIN != 0 || ((B / NMIN) & (B / NMIN - 1)) != 0)
        return fail(ctx, LATER_B200_EINVAL, "block_cols must be 128 * 2^k");
    if (n < B || n % B != 0) return fail(ctx, LATER_B200_EINVAL, "n must be a multiple of block_cols");
    if (m < n || m % 8 != 0) return fail(ctx, LATER_B200_EINVAL, "m must be >= n and a multiple of 8");
    if (lda < m || ldr < n) return fail(ctx, LATER_B200_EINVAL, "leading dimension too small");
    DeviceGuard guard(ctx->device);
    cudaError_t e = guard.error();
    if (e != cudaSuccess) return cuda_fail(ctx, e, "cudaSetDevice");
    if (!ctx->s_in) {
        if ((e = cudaStreamCreateWithFlags(&ctx->s_in, cudaStreamNonBlocking)) != cudaSuccess ||
            (e = cudaStreamCreateWithFlags(&ctx->s_out, cudaStreamNonBlocking)) != cudaSuccess)
            return cuda_fail(ctx, e, "copy streams");
    }
    // device window: m x 3B fp32 + its 3B x 3B R
    const int W = 3 * B;
    const size_t a_bytes = (size_t)m * W * sizeof(float), r_bytes = (size_t)W * W * sizeof(float);
    if (ctx->dA_bytes < a_bytes) {
        if (ctx->dA) cudaFree(ctx->dA);
        ctx->dA = nullptr; ctx->dA_bytes = 0;
        if ((e = cudaMalloc(&ctx->dA, a_bytes)) != cudaSuccess) return cuda_fail(ctx, e, "cudaMalloc window");
        ctx->dA_bytes = a_bytes;
    }
    if (ctx->dR_bytes < r_bytes) {
        if (ctx->dR) cudaFree(ctx->dR);
        ctx->dR = nullptr; ctx->dR_bytes = 0;
        if ((e = cudaMalloc(&ctx->dR, r_bytes)) != cudaSuccess) return cuda_fail(ctx, e, "cudaMalloc R");
        ctx->dR_bytes = r_bytes;
    }
    float* D = ctx->dA;
    float* Rw = ctx->dR;
    int rc = prepare_plan(ctx, m, W, D, m, Rw, W);
    if (rc) return rc;
    auto& p = ctx->plan;
    Recursion rec{};
    rec.ctx = ctx;
    rec.p = &p;
    HalfMatrix qm{p.Qh, p.m, p.n, p.ldh};
    if ((e = make_tensor_map_f16(&rec.q128, qm, 64, 128)) != cudaSuccess ||
        (e = make_tensor_map_f16(&rec.q256, qm, 64, 256)) != cudaSuccess ||
        (e = make_tensor_map_f16(&rec.q64, qm, 64, 64)) != cudaSuccess)
        return cuda_fail(ctx, e, "tensor map encode");
    cudaStream_t st = ctx->stream;
    cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_free[2] = {nullptr, nullptr}, ev_r = nullptr;
    for (auto* ev : {&ev_in[0], &ev_in[1], &ev_free[0], &ev_free[1], &ev_r})
        if ((e = cudaEventCreateWithFlags(ev, cudaEventDisableTiming)) != cudaSuccess) return cuda_fail(ctx, e, "event");
    auto cleanup = [&]() { for (cudaEvent_t ev : {ev_in[0], ev_in[1], ev_free[0], ev_free[1], ev_r}) if (ev) cudaEventDestroy(ev); };
    const size_t col = (size_t)m * sizeof(float);
    auto h2d_block = [&](int slot, int jb, cudaStream_t s) {
        return cudaMemcpy2DAsync(D + (size_t)slot * B * m, col, hA + (size_t)jb * B * lda, (size_t)lda * sizeof(float), col, B,
                                 cudaMemcpyHostToDevice, s);
    };
    if ((e = cudaMemsetAsync(ctx->d_info, 0, kInfoWords * sizeof(int), st)) != cudaSuccess) { cleanup(); return cuda_fail(ctx, e, "info clear"); }
    const int nb = n / B;
    cudaError_t err = cudaSuccess;
    auto ck = [&](cudaError_t x) { if (err == cudaSuccess && x != cudaSuccess) err = x; };
    for (int j = 0; j < nb && err == cudaSuccess && rec.err == cudaSuccess; ++j) {
        ck(h2d_block(2, j, st));                                     // the block to work on
        if (j > 0) {                                                 // first finished block on its way
            // (the copy stream re-reads finished Q blocks from the host: it must stay behind the main
            // stream's copy-out of the newest of them)
            ck(cudaEventRecord(ev_r, st));
            ck(cudaStreamWaitEvent(ctx->s_in, ev_r, 0));
            ck(cudaStreamWaitEvent(ctx->s_in, ev_free[0], 0));
            ck(h2d_block(0, 0, ctx->s_in));
            ck(cudaEventRecord(ev_in[0], ctx->s_in));
        }
        for (int i = 0; i < j && err == cudaSuccess; ++i) {
            const int slot = i & 1;
            if (i + 1 < j) {                                         // prefetch the next Q block into the other slot
                ck(cudaStreamWaitEvent(ctx->s_in, ev_free[slot ^ 1], 0));
                ck(h2d_block(slot ^ 1, i + 1, ctx->s_in));
                ck(cudaEventRecord(ev_in[slot ^ 1], ctx->s_in));
            }
            ck(cudaStreamWaitEvent(st, ev_in[slot], 0));
            rec.cast(slot * B, slot * B + B);                        // fp16 operand of the two products
            rec.gram_update(slot * B, B, 2 * B, B, false, i == 0, i + 1 == j);   // R_ij, A_j -= Q_i R_ij
            ck(cudaEventRecord(ev_free[slot], st));
            // R_ij -> host (ordered behind the products on the main stream; B x B, small)
            ck(cudaMemcpy2DAsync(hR + (size_t)i * B + (size_t)j * B * ldr, (size_t)ldr * sizeof(float),
                                 Rw + (size_t)slot * B + (size_t)2 * B * W, (size_t)W * sizeof(float), (size_t)B * sizeof(float), B,
                                 cudaMemcpyDeviceToHost, st));
        }
        if (j == 0) rec.cast(2 * B, 3 * B);                          // (later blocks: the updates refreshed the shadow)
        rec.qr(2 * B, B);
        ck(cudaMemcpy2DAsync(hR + (size_t)j * B + (size_t)j * B * ldr, (size_t)ldr * sizeof(float),
                             Rw + (size_t)2 * B + (size_t)2 * B * W, (size_t)W * sizeof(float), (size_t)B * sizeof(float), B,
                             cudaMemcpyDeviceToHost, st));
        ck(cudaMemcpy2DAsync(hA + (size_t)j * B * lda, (size_t)lda * sizeof(float), D + (size_t)2 * B * m, col, col, B,
                             cudaMemcpyDeviceToHost, st));
        if (j == 0) { ck(cudaEventRecord(ev_free[0], st)); ck(cudaEventRecord(ev_free[1], st)); }
    }
    if (err == cudaSuccess && rec.err == cudaSuccess)
        ck(cudaMemcpyAsync(ctx->h_info, ctx->d_info, kInfoWords * sizeof(int), cudaMemcpyDeviceToHost, st));
    ck(cudaStreamSynchronize(ctx->s_in));
    ck(cudaStreamSynchronize(st));
    cleanup();
    ctx->plan.valid = false;
    ctx->launches = rec.launches;
    if (rec.err != cudaSuccess) return cuda_fail(ctx, rec.err, "oc_qr enqueue");
    if (err != cudaSuccess) return cuda_fail(ctx, err, "oc_qr");
    if (ctx->h_info[INFO_BAD_COLUMN] != 0 || (ctx->h_info[INFO_FLAGS] & 1))
        return fail(ctx, LATER_B200_ERANK, rank_message(ctx->h_info));
    return 0;
}

// Shared by the two host-input entry points: factor the device matrix (dA, dR) while its columns
// arrive from hA; with hR != nullptr, Q (into hA) and R (into hR) also stream back.
static int rgsqrf_streamed(later_b200_ctx* ctx, int m, int n, float* hA, int hlda, float* hR, int hldr,
                           float* dA, int ldda, float* dR, int lddr) {
    cudaError_t e;
    if (!ctx->s_in) {
        if ((e = cudaStreamCreateWithFlags(&ctx->s_in, cudaStreamNonBlocking)) != cudaSuccess ||
            (e = cudaStreamCreateWithFlags(&ctx->s_out, cudaStreamNonBlocking)) != cudaSuccess)
            return cuda_fail(ctx, e, "copy streams");
    }
    int rc = rgsqrf_prepare(ctx, m, n, dA, ldda, dR, lddr);
    if (rc) return rc;
    auto& p = ctx->plan;
    p.hA = hA; p.hlda = hlda; p.hR = hR; p.hldr = hldr;
    // The copies only overlap (and can only be graph nodes that replay safely) from page-locked
    // memory; with pageable buffers the same sequence is enqueued directly and the runtime stages it.
    cudaPointerAttributes pa{}, pr{};
    const bool pinned = cudaPointerGetAttributes(&pa, hA) == cudaSuccess && pa.type == cudaMemoryTypeHost &&
                        (!hR || (cudaPointerGetAttributes(&pr, hR) == cudaSuccess &&
                                 pr.type == cudaMemoryTypeHost));
    (void)cudaGetLastError();
    const bool graph = ctx->use_graph;
    if (!pinned) ctx->use_graph = false;
    rc = run_stage(ctx, STAGE_HOST);
    ctx->use_graph = graph;
    if (rc) {
        cudaStreamSynchronize(ctx->s_in);
        cudaStreamSynchronize(ctx->s_out);
        cudaStreamSynchronize(ctx->stream);
        (void)cudaGetLastError();
    }
    return rc;   // the side streams have been joined back into the context's stream
}

int later_b200_rgsqrf_host(later_b200_ctx* ctx, int m, int n, float* hA, int lda, float* hR,
                           int ldr) {
    int rc = validate(ctx, m, n, hA, lda, hR, ldr);
    if (rc) return rc;
    DeviceGuard guard(ctx->device);
    cudaError_t e = guard.error();
    if (e != cudaSuccess) return cuda_fail(ctx, e, "cudaSetDevice");
    const size_t a_bytes = (size_t)m * n * sizeof(float), r_bytes = (size_t)n * n * sizeof(float);
    if (ctx->dA_bytes < a_bytes) {
        if (ctx->dA) cudaFree(ctx->dA);
        ctx->dA = nullptr; ctx->dA_bytes = 0;
        if ((e = cudaMalloc(&ctx->dA, a_bytes)) != cudaSuccess) return cuda_fail(ctx, e, "cudaMalloc A");
        ctx->dA_bytes = a_bytes;
    }
    if (ctx->dR_bytes < r_bytes) {
        if (ctx->dR) cudaFree(ctx->dR);
        ctx->dR = nullptr; ctx->dR_bytes = 0;
        if ((e = cudaMalloc(&ctx->dR, r_bytes)) != cudaSuccess) return cuda_fail(ctx, e, "cudaMalloc R");
        ctx->dR_bytes = r_bytes;
    }
    if ((rc = rgsqrf_streamed(ctx, m, n, hA, lda, hR, ldr, ctx->dA, m, ctx->dR, n)) != 0) return rc;
    if ((e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess) return cuda_fail(ctx, e, "sync");
    // blocking call: the status words are in, so breakdown is reported right here
    if (ctx->h_info[INFO_BAD_COLUMN] != 0 || (ctx->h_info[INFO_FLAGS] & 1))
        return fail(ctx, LATER_B200_ERANK, rank_message(ctx->h_info));
    return 0;
}

int later_b200_rgsqrf_stream_in(later_b200_ctx* ctx, int m, int n, const float* hA, int hlda, float* A,
                                int lda, float* R, int ldr) {
    int rc = validate(ctx, m, n, A, lda, R, ldr);
    if (rc) return rc;
    if (!hA || hlda < m) return fail(ctx, LATER_B200_EINVAL, "bad host matrix");
    DeviceGuard guard(ctx->device);
    if (guard.error() != cudaSuccess) return cuda_fail(ctx, guard.error(), "cudaSetDevice");
    return rgsqrf_streamed(ctx, m, n, const_cast<float*>(hA), hlda, nullptr, 0, A, lda, R, ldr);
}

int later_b200_panel_qr(later_b200_ctx* ctx, int m, int n, float* A, int lda, float* R, int ldr) {
    if (!ctx) return LATER_B200_EINVAL;
    if (n != kPanelWidth) return fail(ctx, LATER_B200_EINVAL, "panel width must be 128");
    if (!A || !R || m < n || lda < m || ldr < n) return fail(ctx, LATER_B200_EINVAL, "bad panel arguments");
    DeviceGuard guard(ctx->device);
    cudaError_t e = guard.error();
    if (e != cudaSuccess) return cuda_fail(ctx, e, "cudaSetDevice");
    const size_t bytes = panel_scratch_bytes(m, ctx->num_sms);
    if ((e = ctx->arena.reserve(bytes + 4096)) != cudaSuccess) {
        cuda_fail(ctx, e, "workspace reserve");
        return LATER_B200_ENOMEM;
    }
    ctx->arena.reset();
    ctx->plan.valid = false;
    void* scratch = ctx->arena.alloc(bytes);
    if ((e = cudaMemsetAsync(ctx->d_info, 0, kInfoWords * sizeof(int), ctx->stream)) != cudaSuccess)
        return cuda_fail(ctx, e, "info clear");
    e = panel_qr128(ctx->stream, ctx->num_sms, m, A, lda, R, ldr, nullptr, 0, scratch, false, ctx->opts.panel,
                    ctx->d_info, 0);
    ctx->launches = panel_launch_count(m, ctx->num_sms, A, lda, false, ctx->opts.panel);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "panel_qr128");
    if ((e = cudaMemcpyAsync(ctx->h_info, ctx->d_info, kInfoWords * sizeof(int), cudaMemcpyDeviceToHost,
                             ctx->stream)) != cudaSuccess)
        return cuda_fail(ctx, e, "info read-back");
    return 0;
}

int later_b200_tsqr_apply(later_b200_ctx* ctx, int m, int n, float* Q, int ldq, const float* W,
                          int ldw) {
    if (!ctx) return LATER_B200_EINVAL;
    auto& p = ctx->plan;
    if (!p.valid || p.m != m || p.n != n || p.A != Q || p.lda != ldq)
        return fail(ctx, LATER_B200_ESTATE, "tsqr_apply needs the preceding rgsqrf on the same Q");
    if (!W || ldw < n) return fail(ctx, LATER_B200_EINVAL, "bad W");
    DeviceGuard guard(ctx->device);
    cudaError_t e = guard.error();
    if (e != cudaSuccess) return cuda_fail(ctx, e, "cudaSetDevice");
    const long cast_blocks = std::min<long>(1184, ((long)n * n + 255) / 256);
    cast_matrix_kernel<<<(unsigned)cast_blocks, 256, 0, ctx->stream>>>(W, ldw, n, n, p.Wh, n);
    if ((e = cudaGetLastError()) != cudaSuccess) return cuda_fail(ctx, e, "tsqr_apply cast");
    CUtensorMap q64, wmap;
    HalfMatrix qm{p.Qh, p.m, p.n, p.ldh}, wm{p.Wh, n, n, n};
    const int bn = n >= 256 ? 256 : 128;
    if ((e = make_tensor_map_f16(&q64, qm, 64, 64)) == cudaSuccess &&
        (e = make_tensor_map_f16(&wmap, wm, 64, bn)) == cudaSuccess)
        e = (ldq % 4 == 0 && (reinterpret_cast<uintptr_t>(Q) & 15) == 0)
                ? tc_update_tma(ctx->stream, ctx->num_sms, q64, wmap, bn, 0, m, 0, n, 0, n, Q, m, n, ldq,
                                0, nullptr, 0, false)
                : tc_update(ctx->stream, ctx->num_sms, q64, wmap, bn, 0, m, 0, n, 0, n, Q, ldq, nullptr,
                            0, false);
    ctx->launches = 2;
    if (e != cudaSuccess) return cuda_fail(ctx, e, "tsqr_apply");
    return 0;
}

int later_b200_gemm_gram(later_b200_ctx* ctx, const void* Qh, int q_rows, int q_cols, long ldq,
                         int colA, int Mc, int colB, int Nc, float* C, long ldc, void* Ch,
                         long ldch, int splits) {
    if (!ctx || !Qh || !C) return LATER_B200_EINVAL;
    DeviceGuard guard(ctx->device);
    cudaError_t e = guard.error();
    if (e != cudaSuccess) return cuda_fail(ctx, e, "cudaSetDevice");
    const int bn = Nc >= 256 ? 256 : 128;
    if (splits <= 0) splits = choose_gram_splits(ctx->num_sms, Mc, Nc, bn, q_rows);
    float* part = nullptr;
    if (splits > 1) {
        if ((e = ctx->arena.reserve((size_t)splits * Mc * Nc * sizeof(float) + 4096)) != cudaSuccess) {
            cuda_fail(ctx, e, "workspace reserve");
            return LATER_B200_ENOMEM;
        }
        ctx->arena.reset();
        ctx->plan.valid = false;
        part = static_cast<float*>(ctx->arena.alloc((size_t)splits * Mc * Nc * sizeof(float)));
    }
    CUtensorMap q128, qbn;
    HalfMatrix qm{static_cast<const __half*>(Qh), q_rows, q_cols, ldq};
    if ((e = make_tensor_map_f16(&q128, qm, 64, 128)) != cudaSuccess ||
        (e = make_tensor_map_f16(&qbn, qm, 64, bn)) != cudaSuccess)
        return cuda_fail(ctx, e, "tensor map encode");
    e = tc_gram(ctx->stream, ctx->num_sms, q128, qbn, bn, 0, q_rows, colA, Mc, colB, Nc, C, ldc,
                static_cast<__half*>(Ch), ldch, part, splits);
    ctx->launches = splits > 1 ? 2 : 1;
    if (e != cudaSuccess) return cuda_fail(ctx, e, "tc_gram");
    return 0;
}

int later_b200_gemm_update(later_b200_ctx* ctx, const void* Qh, int q_rows, int q_cols, long ldq,
                           int colA, int K, const void* Bh, long ldb, int Nc, float* C, long ldc,
                           void* Ch, long ldch, int subtract) {
    if (!ctx || !Qh || !Bh || !C) return LATER_B200_EINVAL;
    DeviceGuard guard(ctx->device);
    cudaError_t e = guard.error();
    if (e != cudaSuccess) return cuda_fail(ctx, e, "cudaSetDevice");
    int bn = (Nc >= 256 && K >= 2048) ? 256 : 128;
    const int variant = ctx->opts.update_variant;   // diagnostics only
    if (variant == 1 || variant == 2) bn = Nc >= 256 ? 256 : 128;
    if (variant == 3) bn = 128;
    CUtensorMap q64, bmap;
    HalfMatrix qm{static_cast<const __half*>(Qh), q_rows, q_cols, ldq};
    HalfMatrix bm{static_cast<const __half*>(Bh), K, Nc, ldb};
    if ((e = make_tensor_map_f16(&q64, qm, 64, 64)) != cudaSuccess ||
        (e = make_tensor_map_f16(&bmap, bm, 64, bn)) != cudaSuccess)
        return cuda_fail(ctx, e, "tensor map encode");
    const bool tma_ok = ldc % 4 == 0 && (reinterpret_cast<uintptr_t>(C) & 15) == 0 &&
                        (subtract == 0 ? Ch == nullptr
                                       : (Ch && ldch % 8 == 0 && (reinterpret_cast<uintptr_t>(Ch) & 15) == 0)) &&
                        Nc % bn == 0;
    if (tma_ok && variant != 1)
        e = tc_update_tma(ctx->stream, ctx->num_sms, q64, bmap, bn, 0, q_rows, colA, K, 0, Nc, C, q_rows,
                          Nc, ldc, 0, static_cast<__half*>(Ch), ldch, subtract != 0);
    else
        e = tc_update(ctx->stream, ctx->num_sms, q64, bmap, bn, 0, q_rows, colA, K, 0, Nc, C, ldc,
                      static_cast<__half*>(Ch), ldch, subtract != 0);
    ctx->launches = 1;
    if (e != cudaSuccess) return cuda_fail(ctx, e, "tc_update");
    return 0;
}

}  // extern "C"
