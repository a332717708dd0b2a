// optim.cuh — element arithmetic of the student update, shared by the single-GPU multi-tensor step
// (optim.cu) and the data-parallel sharded step over peer memory (dp.cu).  Same op order as torch's
// single-tensor Adam / SGD (torch/optim/adam.py::_single_tensor_adam, sgd.py::_single_tensor_sgd); the EMA
// keeps the three fp32 roundings of the reference's  p.mul_(a); p.add_(s*(1-a))  (utils.py:24-25).
#pragma once

#include "common.cuh"

namespace udape {

constexpr int kOptThreads = 256;
constexpr int kOptUnroll = 2;

struct OptScalars {
    float inv_scale;   // 1 / grad_scale (1 when no scaler)
    float lr;
    float w1;          // Adam: 1 - beta1          SGD: momentum
    float beta2;       // Adam: beta2              SGD: 1 - dampening
    float w2;          // Adam: 1 - beta2
    float eps;
    float wd;
    float neg_step;    // Adam: -(lr / bias_correction1)   SGD: -lr
    float bc2_sqrt;    // Adam: sqrt(bias_correction2)
    int skip;          // found_inf != 0: leave the student alone
};

__device__ __forceinline__ float ema_fold(float t, float s, float a, float b) {
    return __fadd_rn(__fmul_rn(t, a), __fmul_rn(s, b));
}

__device__ __forceinline__ void adam_elem(float& p, float g, float& m, float& v, const OptScalars& c) {
    g *= c.inv_scale;                                   // GradScaler.unscale_: grad.mul_(inv_scale)
    if (c.wd != 0.0f) g = fmaf(c.wd, p, g);             // grad.add(param, alpha=weight_decay)
    m = fmaf(c.w1, g - m, m);                           // exp_avg.lerp_(grad, 1 - beta1)
    v = fmaf(c.w2 * g, g, v * c.beta2);                 // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    const float denom = sqrtf(v) / c.bc2_sqrt + c.eps;  // (exp_avg_sq.sqrt() / bc2_sqrt).add_(eps)
    p = fmaf(c.neg_step, m / denom, p);                 // param.addcdiv_(exp_avg, denom, value=-step_size)
}

template <bool NESTEROV>
__device__ __forceinline__ void sgd_elem(float& p, float g, float& buf, const OptScalars& c, bool momentum,
                                         bool first) {
    g *= c.inv_scale;
    if (c.wd != 0.0f) g = fmaf(c.wd, p, g);
    if (momentum) {
        // first update of THIS parameter: buf = clone(grad); else buf.mul_(momentum).add_(grad, alpha=1-dampening)
        buf = first ? g : fmaf(c.beta2, g, buf * c.w1);
        g = NESTEROV ? fmaf(c.w1, buf, g) : buf;            // grad.add(buf, alpha=momentum)
    }
    p = fmaf(-c.lr, g, p);                                  // param.add_(grad, alpha=-lr)
}

// Scalars of one update from the hyper-parameters and the device-resident step count / lr / loss scale.
// ALGO 0: Adam, otherwise SGD.  Bias corrections and the step size are formed in double and rounded once.
template <int ALGO>
__device__ __forceinline__ OptScalars make_opt_scalars(const udape_opt_hyper& h, const float* lr_dev,
                                                       const float* grad_scale, const float* found_inf,
                                                       const int32_t* step_dev) {
    OptScalars s;
    const double lr = lr_dev ? static_cast<double>(*lr_dev) : h.lr;
    s.skip = (found_inf && *found_inf != 0.0f) ? 1 : 0;
    const int step = (step_dev ? *step_dev : h.step - 1) + 1;  // number of THIS update, 1-based
    // GradScaler: inv_scale = scale.double().reciprocal().float()
    s.inv_scale = grad_scale ? static_cast<float>(1.0 / static_cast<double>(*grad_scale)) : 1.0f;
    s.lr = static_cast<float>(lr);
    s.eps = static_cast<float>(h.eps);
    s.wd = static_cast<float>(h.weight_decay);
    if (ALGO == 0) {
        s.w1 = static_cast<float>(1.0 - h.beta1);
        s.beta2 = static_cast<float>(h.beta2);
        s.w2 = static_cast<float>(1.0 - h.beta2);
        const double bc1 = 1.0 - pow(h.beta1, static_cast<double>(step));
        const double bc2 = 1.0 - pow(h.beta2, static_cast<double>(step));
        s.neg_step = static_cast<float>(-(lr / bc1));
        s.bc2_sqrt = static_cast<float>(sqrt(bc2));
    } else {
        s.w1 = static_cast<float>(h.beta1);          // momentum
        s.beta2 = static_cast<float>(1.0 - h.beta2);  // 1 - dampening
        s.w2 = 0.0f;
        s.neg_step = -s.lr;
        s.bc2_sqrt = 1.0f;
    }
    return s;
}

}  // namespace udape
