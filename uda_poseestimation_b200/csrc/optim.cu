// optim.cu — the student update and the teacher EMA in one multi-tensor pass.
//
// Replaces the tail of the reference's train step (train_human.py:436-440):
//     scaler.scale(loss_all).backward(); scaler.step(stu_optimizer); tea_optimizer.step(); scaler.update()
// i.e. GradScaler.unscale_ (read + write of every gradient, plus a non-finite check and a host
// `.item()` sync), torch.optim.Adam / SGD(momentum=0.9, nesterov) (train_human.py:136-139; a dozen
// foreach passes over param / grad / exp_avg / exp_avg_sq) and OldWeightEMA.step (utils.py:21-25,
// 969 launches).  Here:
//   * grad_check_kernel   — one READ-ONLY pass over the gradients -> found_inf (device float; no
//                           host sync, no unscaled copy written back);
//   * student_step_kernel — one pass that unscales the gradient in registers, applies Adam or SGD to
//                           the student, and folds the updated student into the teacher EMA:
//                           Adam 5 reads + 4 writes per element (36 B) instead of the reference's
//                           ~20 passes; SGD 4 reads + 3 writes.
// When found_inf is set the student (and its optimizer state and step count) is left untouched and
// only the EMA runs — exactly what scaler.step() + tea_optimizer.step() do in the reference.
//
// Arithmetic follows torch's single-tensor Adam/SGD op order (torch/optim/adam.py::_single_tensor_adam,
// sgd.py::_single_tensor_sgd): bias corrections and step size are formed in double precision and
// rounded to float once; the EMA keeps the three fp32 roundings of  p.mul_(a); p.add_(s*(1-a)).
#include "optim.cuh"

namespace udape {

// ALGO 0: Adam, 1: SGD, 2: SGD + Nesterov
template <int ALGO>
__global__ void __launch_bounds__(kOptThreads)
student_step_kernel(const udape_opt_chunk* __restrict__ chunks, udape_opt_hyper h,
                    const float* __restrict__ lr_dev, const float* __restrict__ grad_scale,
                    const float* __restrict__ found_inf, int32_t* __restrict__ step_dev, int advance_step,
                    int32_t* __restrict__ fresh_flags, int n_fresh, uint32_t* __restrict__ ticket) {
    __shared__ OptScalars sc;
    const udape_opt_chunk c = chunks[blockIdx.x];
    if (threadIdx.x == 0) {
        sc = make_opt_scalars<ALGO>(h, lr_dev, grad_scale, found_inf, step_dev);
    }
    __syncthreads();
    const OptScalars s = sc;
    const float ea = h.ema_a, eb = h.ema_b;

    float* __restrict__ p = static_cast<float*>(c.param);
    const float* __restrict__ g = static_cast<const float*>(c.grad);
    float* __restrict__ m = static_cast<float*>(c.state1);
    float* __restrict__ v = static_cast<float*>(c.state2);
    float* __restrict__ t = static_cast<float*>(c.ema);
    const int n = static_cast<int>(c.numel);
    const bool upd = !s.skip && g != nullptr;            // torch skips parameters whose grad is None
    const bool has_m = m != nullptr;                     // SGD with momentum == 0 keeps no buffer
    // torch/optim/sgd.py: `if buf is None: buf = torch.clone(grad)` — per parameter, the first time it is updated
    const bool first = ALGO != 0 && c.fresh != nullptr && *c.fresh != 0;
    const bool vec_ok = aligned16(p) && (!upd || (aligned16(g) && (!has_m || aligned16(m)) &&
                                                  (ALGO != 0 || aligned16(v)))) && (!t || aligned16(t));
    int done = 0;
    if (vec_ok) {
        const int nvec = n >> 2;
        for (int base = 0; base < nvec; base += kOptThreads * kOptUnroll) {
            uint4 pv[kOptUnroll], gv[kOptUnroll], mv[kOptUnroll], vv[kOptUnroll], tv[kOptUnroll];
#pragma unroll
            for (int u = 0; u < kOptUnroll; ++u) {   // every load of the thread is issued before the first use
                const int i = base + u * kOptThreads + threadIdx.x;
                if (i < nvec) {
                    pv[u] = ldg_cached(reinterpret_cast<const uint4*>(p) + i);
                    if (upd) {
                        gv[u] = ldg_stream(reinterpret_cast<const uint4*>(g) + i);
                        if (has_m) mv[u] = ldg_cached(reinterpret_cast<const uint4*>(m) + i);
                        if (ALGO == 0) vv[u] = ldg_cached(reinterpret_cast<const uint4*>(v) + i);
                    }
                    if (t) tv[u] = ldg_cached(reinterpret_cast<const uint4*>(t) + i);
                }
            }
#pragma unroll
            for (int u = 0; u < kOptUnroll; ++u) {
                const int i = base + u * kOptThreads + threadIdx.x;
                if (i < nvec) {
                    float fp[4], fg[4], fm[4], fv[4], ft[4];
                    unpack16<float>(pv[u], fp);
                    if (upd) {
                        unpack16<float>(gv[u], fg);
                        if (has_m) unpack16<float>(mv[u], fm);
                        if (ALGO == 0) unpack16<float>(vv[u], fv);
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            if (ALGO == 0) adam_elem(fp[e], fg[e], fm[e], fv[e], s);
                            else sgd_elem<ALGO == 2>(fp[e], fg[e], fm[e], s, has_m, first);
                        }
                        stg_plain(reinterpret_cast<uint4*>(p) + i, pack16<float>(fp));
                        if (has_m) stg_plain(reinterpret_cast<uint4*>(m) + i, pack16<float>(fm));
                        if (ALGO == 0) stg_plain(reinterpret_cast<uint4*>(v) + i, pack16<float>(fv));
                    }
                    if (t) {
                        unpack16<float>(tv[u], ft);
#pragma unroll
                        for (int e = 0; e < 4; ++e) ft[e] = ema_fold(ft[e], fp[e], ea, eb);
                        stg_plain(reinterpret_cast<uint4*>(t) + i, pack16<float>(ft));
                    }
                }
            }
        }
        done = nvec << 2;
    }
    for (int i = done + threadIdx.x; i < n; i += kOptThreads) {
        float fp = p[i];
        if (upd) {
            float fm = has_m ? m[i] : 0.0f;
            if (ALGO == 0) {
                float fv = v[i];
                adam_elem(fp, g[i], fm, fv, s);
                v[i] = fv;
            } else {
                sgd_elem<ALGO == 2>(fp, g[i], fm, s, has_m, first);
            }
            p[i] = fp;
            if (has_m) m[i] = fm;
        }
        if (t) t[i] = ema_fold(t[i], fp, ea, eb);
    }
    // the step counter advances once per launch, after every CTA has read it, and only when the
    // update was applied (torch: state['step'] is not touched when scaler.step() skips)
    // (ticket == NULL: the counter is only read — the launches of all but the last param group of a step)
    if (ticket) {
        if (last_block_done(ticket, gridDim.x) && !s.skip) {
            if (threadIdx.x == 0 && step_dev && advance_step) *step_dev = *step_dev + 1;
            for (int i = threadIdx.x; i < n_fresh; i += kOptThreads) fresh_flags[i] = 0;
        }
    }
}

// found_inf := any(!isfinite(grad))  — torch._amp_foreach_non_finite_check_and_unscale_ without the
// write-back.  ws[0] is a self-resetting ticket, ws[1] the OR word (left zero).
constexpr int kCheckChunksPerCta = 4;   // one barrier + one ticket per 4 chunks (64 KB of gradients)

__global__ void __launch_bounds__(kOptThreads)
grad_check_kernel(const udape_opt_chunk* __restrict__ chunks, int64_t n_chunks, float* __restrict__ found_inf,
                  uint32_t* __restrict__ ws) {
    // a float is non-finite iff its exponent field is all ones: fold with AND over (x & 0x7f800000) == 0x7f800000
    uint32_t bad = 0;
#pragma unroll 1
    for (int k = 0; k < kCheckChunksPerCta; ++k) {
    const int64_t ci = static_cast<int64_t>(blockIdx.x) * kCheckChunksPerCta + k;
    if (ci >= n_chunks) break;
    const udape_opt_chunk c = chunks[ci];
    const float* __restrict__ g = static_cast<const float*>(c.grad);
    const int n = static_cast<int>(c.numel);
    if (g) {
        int done = 0;
        if (aligned16(g)) {
            const int nvec = n >> 2;
            const uint4* g4 = reinterpret_cast<const uint4*>(g);
            for (int base = 0; base < nvec; base += kOptThreads * 4) {
                uint4 x[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int i = base + u * kOptThreads + threadIdx.x;
                    x[u] = i < nvec ? ldg_stream(g4 + i) : make_uint4(0, 0, 0, 0);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    bad |= ((x[u].x & 0x7f800000u) == 0x7f800000u) | ((x[u].y & 0x7f800000u) == 0x7f800000u) |
                           ((x[u].z & 0x7f800000u) == 0x7f800000u) | ((x[u].w & 0x7f800000u) == 0x7f800000u);
                }
            }
            done = nvec << 2;
        }
        for (int i = done + threadIdx.x; i < n; i += kOptThreads)
            bad |= (__float_as_uint(g[i]) & 0x7f800000u) == 0x7f800000u;
    }
    }
    const int any = __syncthreads_or(static_cast<int>(bad));
    if (threadIdx.x == 0 && any) atomicOr(ws + 1, 1u);
    if (last_block_done(ws, gridDim.x) && threadIdx.x == 0) {
        const uint32_t f = *reinterpret_cast<volatile uint32_t*>(ws + 1);
        *found_inf = f ? 1.0f : 0.0f;
        ws[1] = 0u;
    }
}

}  // namespace udape

using namespace udape;

extern "C" int64_t udape_opt_plan(void* const* param, const void* const* grad, void* const* state1,
                                  void* const* state2, void* const* ema, const int32_t* const* fresh,
                                  const int64_t* numel, int64_t n_tensors, int64_t chunk_elems,
                                  udape_opt_chunk* out, int64_t capacity) {
    if (!param || !numel) return fail(UDAPE_ERR_NULL, "udape_opt_plan: NULL table");
    if (n_tensors < 0 || chunk_elems <= 0 || chunk_elems >= (1ll << 31) || (chunk_elems % 16) != 0)
        return fail(UDAPE_ERR_ARG, "udape_opt_plan: chunk_elems must be a positive multiple of 16 below 2^31");
    int64_t n = 0;
    for (int64_t t = 0; t < n_tensors; ++t) {
        if (numel[t] < 0) return fail(UDAPE_ERR_SHAPE, "udape_opt_plan: tensor %lld has negative numel", (long long)t);
        if (numel[t] > 0 && !param[t])
            return fail(UDAPE_ERR_NULL, "udape_opt_plan: tensor %lld has a NULL parameter pointer", (long long)t);
        for (int64_t off = 0; off < numel[t]; off += chunk_elems, ++n) {
            if (out && n < capacity) {
                const int64_t rem = numel[t] - off;
                auto at = [off](const void* base) -> void* {
                    return base ? const_cast<char*>(static_cast<const char*>(base)) + off * 4 : nullptr;
                };
                out[n].param = at(param[t]);
                out[n].grad = grad ? at(grad[t]) : nullptr;
                out[n].state1 = state1 ? at(state1[t]) : nullptr;
                out[n].state2 = state2 ? at(state2[t]) : nullptr;
                out[n].ema = ema ? at(ema[t]) : nullptr;
                out[n].fresh = fresh ? fresh[t] : nullptr;
                out[n].numel = rem < chunk_elems ? rem : chunk_elems;
            }
        }
    }
    return n;
}

extern "C" int udape_grad_check(const udape_opt_chunk* chunks_dev, int64_t n_chunks, float* found_inf,
                                uint32_t* ws, void* stream) {
    UDAPE_REQUIRE(found_inf && ws, UDAPE_ERR_NULL, "udape_grad_check: found_inf / ws is NULL");
    UDAPE_REQUIRE(n_chunks >= 0 && n_chunks < (1ll << 31), UDAPE_ERR_SHAPE, "udape_grad_check: bad n_chunks=%lld", (long long)n_chunks);
    cudaStream_t st = as_stream(stream);
    if (n_chunks == 0) {
        cudaError_t e = cudaMemsetAsync(found_inf, 0, sizeof(float), st);
        return e == cudaSuccess ? UDAPE_OK : fail(static_cast<int>(e), "udape_grad_check: %s", cudaGetErrorString(e));
    }
    UDAPE_REQUIRE(chunks_dev, UDAPE_ERR_NULL, "udape_grad_check: chunk table is NULL");
    const unsigned grid = static_cast<unsigned>((n_chunks + kCheckChunksPerCta - 1) / kCheckChunksPerCta);
    grad_check_kernel<<<grid, kOptThreads, 0, st>>>(chunks_dev, n_chunks, found_inf, ws);
    return check_launch("udape_grad_check");
}

extern "C" int udape_student_step(const udape_opt_chunk* chunks_dev, int64_t n_chunks, int algo,
                                  const udape_opt_hyper* hyper, const float* lr_dev, const float* grad_scale,
                                  const float* found_inf, int32_t* step_dev, int advance_step,
                                  int32_t* fresh_flags, int64_t n_fresh, uint32_t* ticket, void* stream) {
    if (n_chunks == 0) return UDAPE_OK;
    UDAPE_REQUIRE(chunks_dev && hyper, UDAPE_ERR_NULL, "udape_student_step: chunk table / hyper is NULL");
    UDAPE_REQUIRE(n_chunks > 0 && n_chunks < (1ll << 31), UDAPE_ERR_SHAPE, "udape_student_step: bad n_chunks=%lld", (long long)n_chunks);
    UDAPE_REQUIRE(algo == UDAPE_OPT_ADAM || algo == UDAPE_OPT_SGD, UDAPE_ERR_ARG, "udape_student_step: algo must be UDAPE_OPT_ADAM or UDAPE_OPT_SGD");
    UDAPE_REQUIRE(step_dev || hyper->step >= 1, UDAPE_ERR_ARG, "udape_student_step: hyper.step is 1-based (got %d)", (int)hyper->step);
    if (algo == UDAPE_OPT_ADAM)
        UDAPE_REQUIRE(hyper->beta1 >= 0.0 && hyper->beta1 < 1.0 && hyper->beta2 >= 0.0 && hyper->beta2 < 1.0 && hyper->eps >= 0.0,
                      UDAPE_ERR_ARG, "udape_student_step: Adam needs 0 <= beta < 1 and eps >= 0");
    UDAPE_REQUIRE(!(hyper->nesterov && algo == UDAPE_OPT_SGD && (hyper->beta1 <= 0.0 || hyper->beta2 != 0.0)),
                  UDAPE_ERR_ARG, "udape_student_step: Nesterov momentum requires a momentum and zero dampening");
    UDAPE_REQUIRE(n_fresh >= 0 && n_fresh < (1ll << 31) && (n_fresh == 0 || (fresh_flags && ticket)), UDAPE_ERR_ARG,
                  "udape_student_step: n_fresh=%lld needs fresh_flags and a ticket", (long long)n_fresh);
    cudaStream_t st = as_stream(stream);
    const unsigned grid = static_cast<unsigned>(n_chunks);
    const int nf = static_cast<int>(n_fresh);
    if (algo == UDAPE_OPT_ADAM)
        student_step_kernel<0><<<grid, kOptThreads, 0, st>>>(chunks_dev, *hyper, lr_dev, grad_scale, found_inf, step_dev, advance_step, fresh_flags, nf, ticket);
    else if (hyper->nesterov)
        student_step_kernel<2><<<grid, kOptThreads, 0, st>>>(chunks_dev, *hyper, lr_dev, grad_scale, found_inf, step_dev, advance_step, fresh_flags, nf, ticket);
    else
        student_step_kernel<1><<<grid, kOptThreads, 0, st>>>(chunks_dev, *hyper, lr_dev, grad_scale, found_inf, step_dev, advance_step, fresh_flags, nf, ticket);
    return check_launch("udape_student_step");
}
