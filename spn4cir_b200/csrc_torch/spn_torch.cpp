// spn4cir_torch: the PyTorch extension over the C ABI (include/spn_b200.h) -- the reference call shape
//     loss = model.infonce_loss(q, bank, labels, tau); scaler.scale(loss).backward()
// (clip4cir/models_negplus.py:150-154 driven by clip4cir/train_negplus.py:110-121) as torch.library operators with a C++
// autograd node, so that a training step costs a few tens of microseconds of host time instead of the Python
// autograd.Function hand-off (~140-200 us per step in round 1, longer than the GPU step itself).
//
// torch is plumbing here: tensors own the device memory, c10 supplies the current stream, autograd records one node.
// Every arithmetic step runs in libspn_b200's kernels through the same extern "C" entry points the ctypes binding uses
// (spn_bank_ce_step, spn_query_assemble, spn_query_assemble_bwd); no torch type crosses that boundary.
//
// Operators (namespace spn4cir):
//   bank_ce_step(q, bank, labels, maxnorm, colsum?, inv_tau, eps, T, want_ddot) -> (loss, dq_unit, dlogit_dot)
//   bank_infonce(q, tau?, bank, labels, maxnorm, colsum?, inv_tau, eps, T) -> loss        [differentiable: q, tau]
//   bank_large_step_loss(refer_bank, refer_idx, text, tau?, bank, labels, maxnorm, colsum?, inv_tau, eps, T) -> loss
//                                                                                   [differentiable: text, tau]
#include <ATen/ATen.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>
#include <torch/autograd.h>
#include <torch/library.h>

#include <tuple>

#include "../../include/spn_b200.h"

namespace {

using at::Tensor;
using torch::autograd::AutogradContext;
using torch::autograd::variable_list;

int spn_dtype_of(const Tensor& t) {
  switch (t.scalar_type()) {
    case at::kFloat: return SPN_F32;
    case at::kHalf: return SPN_F16;
    case at::kBFloat16: return SPN_BF16;
    default: TORCH_CHECK(false, "spn4cir: unsupported query dtype ", t.scalar_type(), " (float32, float16 or bfloat16)");
  }
}

void check(int rc, const char* what) { TORCH_CHECK(rc == SPN_OK, what, " failed (", rc, "): ", spn_last_error()); }

struct StepOut {
  Tensor loss, dq, ddot;
};

// One spn_bank_ce_step call: prologue -> fused sweep -> finalize.  q: [B, D] fp32/fp16/bf16 (cast to bf16 inside),
// bank: [N * T, D] bf16 resident rows, labels int64 [B], maxnorm fp32 [1], colsum fp32 [D] (label smoothing only).
StepOut step_impl(const Tensor& q_in, const Tensor& bank, const Tensor& labels_in, const Tensor& maxnorm,
                  const c10::optional<Tensor>& colsum, double inv_tau, double eps, int64_t T, bool want_ddot) {
  TORCH_CHECK(q_in.dim() == 2 && bank.dim() == 2 && q_in.size(1) == bank.size(1), "spn4cir: q [B, D] and bank [N*T, D] expected, got ",
              q_in.dim(), "-d q and ", bank.dim(), "-d bank with last sizes ", q_in.dim() ? q_in.size(-1) : 0, " / ",
              bank.dim() ? bank.size(-1) : 0);
  TORCH_CHECK(q_in.is_cuda() && bank.is_cuda() && labels_in.is_cuda(), "spn4cir: CUDA tensors required (no CPU fallback)");
  TORCH_CHECK(q_in.size(0) > 0, "spn4cir: empty batch (B = 0)");
  TORCH_CHECK(bank.scalar_type() == at::kBFloat16 && bank.is_contiguous(), "spn4cir: bank must be contiguous bf16 rows");
  TORCH_CHECK(T >= 1 && bank.size(0) % T == 0, "spn4cir: bank rows not a multiple of T");
  TORCH_CHECK(eps == 0.0 || colsum.has_value(), "spn4cir: label smoothing needs the bank column sums");
  const c10::cuda::CUDAGuard guard(q_in.device());
  Tensor q = q_in.detach();
  if (q.scalar_type() != at::kFloat && q.scalar_type() != at::kHalf && q.scalar_type() != at::kBFloat16) q = q.to(at::kFloat);
  q = q.contiguous();
  Tensor labels = labels_in.scalar_type() == at::kLong ? labels_in.contiguous() : labels_in.to(at::kLong).contiguous();
  const int B = int(q.size(0)), D = int(q.size(1));
  const int64_t N = bank.size(0) / T;
  size_t ws_bytes = 0;
  check(spn_workspace_bytes(SPN_OP_BANK_CE_STEP, B, N, D, int(T), 0, &ws_bytes), "spn_workspace_bytes");
  const auto f32 = q.options().dtype(at::kFloat);
  Tensor ws = at::empty({int64_t(ws_bytes)}, q.options().dtype(at::kByte));
  Tensor aux = at::empty({3, B}, f32);  // lse, loss per query, dlogit_dot
  Tensor loss = at::empty({}, f32);
  Tensor dq = at::empty({B, D}, f32);
  cudaStream_t stream = c10::cuda::getCurrentCUDAStream(q.device().index()).stream();
  check(spn_bank_ce_step(q.data_ptr(), spn_dtype_of(q), bank.data_ptr(), labels.data_ptr<int64_t>(), float(inv_tau),
                         maxnorm.data_ptr<float>(), colsum.has_value() && eps != 0.0 ? colsum->data_ptr<float>() : nullptr,
                         float(eps), B, N, D, int(T), aux[0].data_ptr<float>(), aux[1].data_ptr<float>(),
                         loss.data_ptr<float>(), dq.data_ptr<float>(), want_ddot ? aux[2].data_ptr<float>() : nullptr,
                         ws.data_ptr(), ws_bytes, stream),
        "spn_bank_ce_step");
  return {loss, dq, want_ddot ? aux[2] : Tensor()};
}

std::tuple<Tensor, Tensor, Tensor> bank_ce_step(const Tensor& q, const Tensor& bank, const Tensor& labels,
                                                const Tensor& maxnorm, const c10::optional<Tensor>& colsum, double inv_tau,
                                                double eps, int64_t T, bool want_ddot) {
  StepOut o = step_impl(q, bank, labels, maxnorm, colsum, inv_tau, eps, T, want_ddot);
  return {o.loss, o.dq, o.ddot.defined() ? o.ddot : at::empty({0}, o.loss.options())};
}

// ------------------------------------------------------------------------------------------------ autograd nodes
// forward produces the gradient for an upstream gradient of 1 (one sweep of the bank); backward only scales it.
struct BankInfoNCE : public torch::autograd::Function<BankInfoNCE> {
  static Tensor forward(AutogradContext* ctx, const Tensor& q, const c10::optional<Tensor>& tau_t, const Tensor& bank,
                        const Tensor& labels, const Tensor& maxnorm, const c10::optional<Tensor>& colsum, double inv_tau,
                        double eps, int64_t T) {
    const bool tau_grad = tau_t.has_value() && tau_t->defined() && tau_t->requires_grad();
    at::AutoDispatchBelowADInplaceOrView no_ad;
    StepOut o = step_impl(q, bank, labels, maxnorm, colsum, inv_tau, eps, T, tau_grad);
    ctx->save_for_backward({o.dq, tau_grad ? o.ddot : Tensor()});
    ctx->saved_data["inv_tau"] = inv_tau;
    ctx->saved_data["B"] = q.size(0);
    ctx->saved_data["tau_grad"] = tau_grad;
    ctx->saved_data["q_dtype"] = int64_t(q.scalar_type());
    return o.loss;
  }

  static variable_list backward(AutogradContext* ctx, variable_list grads) {
    const auto saved = ctx->get_saved_variables();
    const Tensor g = grads[0].to(at::kFloat);
    Tensor dq = saved[0] * g;  // the upstream gradient is a device scalar (GradScaler): no host sync
    const auto q_dtype = at::ScalarType(ctx->saved_data["q_dtype"].toInt());
    if (q_dtype != at::kFloat) dq = dq.to(q_dtype);
    Tensor dtau;
    if (ctx->saved_data["tau_grad"].toBool()) {
      // d loss / d tau = -(g / B) * inv_tau * sum_b sum_n dlogit[b, n] s[b, n]
      dtau = saved[1].sum() * g * (-(ctx->saved_data["inv_tau"].toDouble() / double(ctx->saved_data["B"].toInt())));
    }
    return {dq, dtau, Tensor(), Tensor(), Tensor(), Tensor(), Tensor(), Tensor(), Tensor()};
  }
};

Tensor bank_infonce(const Tensor& q, const c10::optional<Tensor>& tau_t, const Tensor& bank, const Tensor& labels,
                    const Tensor& maxnorm, const c10::optional<Tensor>& colsum, double inv_tau, double eps, int64_t T) {
  return BankInfoNCE::apply(q, tau_t, bank, labels, maxnorm, colsum, inv_tau, eps, T);
}

// bank_large_step (clip4cir/models_negplus.py:130-142) fused: q = normalize(refer_bank[idx] + text) is assembled by
// spn_query_assemble (bf16 for the sweep, fp32 + 1/|x| kept for the backward), the loss step runs on it, and the
// backward goes through the normalisation in one kernel (spn_query_assemble_bwd).  refer_bank is detached (:133-135).
struct BankLargeStep : public torch::autograd::Function<BankLargeStep> {
  static Tensor forward(AutogradContext* ctx, const Tensor& refer_bank, const Tensor& refer_idx, const Tensor& text,
                        const c10::optional<Tensor>& tau_t, const Tensor& bank, const Tensor& labels, const Tensor& maxnorm,
                        const c10::optional<Tensor>& colsum, double inv_tau, double eps, int64_t T) {
    const bool tau_grad = tau_t.has_value() && tau_t->defined() && tau_t->requires_grad();
    at::AutoDispatchBelowADInplaceOrView no_ad;
    TORCH_CHECK(refer_bank.is_cuda() && refer_bank.dim() == 2 && text.dim() == 2 && refer_bank.size(1) == text.size(1),
                "spn4cir: refer_bank [R, D] and text [B, D] expected on CUDA");
    const c10::cuda::CUDAGuard guard(text.device());
    Tensor rb = refer_bank.detach().to(at::kFloat).contiguous();
    Tensor tx = text.detach().to(at::kFloat).contiguous();
    Tensor idx = refer_idx.to(text.device(), at::kLong).contiguous();
    const int B = int(tx.size(0)), D = int(tx.size(1));
    TORCH_CHECK(idx.numel() == B, "spn4cir: one reference index per query expected");
    Tensor q16 = at::empty({B, D}, tx.options().dtype(at::kBFloat16));
    Tensor q32 = at::empty({B, D}, tx.options());
    Tensor inv = at::empty({B}, tx.options());
    cudaStream_t stream = c10::cuda::getCurrentCUDAStream(tx.device().index()).stream();
    check(spn_query_assemble(rb.data_ptr<float>(), idx.data_ptr<int64_t>(), tx.data_ptr<float>(), B, D, rb.size(0),
                             q16.data_ptr(), q32.data_ptr<float>(), inv.data_ptr<float>(), stream),
          "spn_query_assemble");
    StepOut o = step_impl(q16, bank, labels, maxnorm, colsum, inv_tau, eps, T, tau_grad);
    ctx->save_for_backward({o.dq, q32, inv, tau_grad ? o.ddot : Tensor()});
    ctx->saved_data["inv_tau"] = inv_tau;
    ctx->saved_data["B"] = int64_t(B);
    ctx->saved_data["tau_grad"] = tau_grad;
    ctx->saved_data["text_dtype"] = int64_t(text.scalar_type());
    return o.loss;
  }

  static variable_list backward(AutogradContext* ctx, variable_list grads) {
    const auto saved = ctx->get_saved_variables();
    const Tensor g = grads[0].to(at::kFloat).contiguous();
    const Tensor &dq = saved[0], &q32 = saved[1], &inv = saved[2];
    const c10::cuda::CUDAGuard guard(dq.device());
    Tensor dtext = at::empty_like(dq);
    cudaStream_t stream = c10::cuda::getCurrentCUDAStream(dq.device().index()).stream();
    check(spn_query_assemble_bwd(dq.data_ptr<float>(), q32.data_ptr<float>(), inv.data_ptr<float>(), g.data_ptr<float>(),
                                 int(dq.size(0)), int(dq.size(1)), dtext.data_ptr<float>(), stream),
          "spn_query_assemble_bwd");
    const auto t_dtype = at::ScalarType(ctx->saved_data["text_dtype"].toInt());
    if (t_dtype != at::kFloat) dtext = dtext.to(t_dtype);
    Tensor dtau;
    if (ctx->saved_data["tau_grad"].toBool())
      dtau = saved[3].sum() * g * (-(ctx->saved_data["inv_tau"].toDouble() / double(ctx->saved_data["B"].toInt())));
    return {Tensor(), Tensor(), dtext, dtau, Tensor(), Tensor(), Tensor(), Tensor(), Tensor(), Tensor(), Tensor()};
  }
};

Tensor bank_large_step_loss(const Tensor& refer_bank, const Tensor& refer_idx, const Tensor& text,
                            const c10::optional<Tensor>& tau_t, const Tensor& bank, const Tensor& labels,
                            const Tensor& maxnorm, const c10::optional<Tensor>& colsum, double inv_tau, double eps, int64_t T) {
  return BankLargeStep::apply(refer_bank, refer_idx, text, tau_t, bank, labels, maxnorm, colsum, inv_tau, eps, T);
}

}  // namespace

TORCH_LIBRARY(spn4cir, m) {
  m.def("bank_ce_step(Tensor q, Tensor bank, Tensor labels, Tensor maxnorm, Tensor? colsum, float inv_tau, float eps, int T, "
        "bool want_ddot) -> (Tensor, Tensor, Tensor)", &bank_ce_step);
  m.def("bank_infonce(Tensor q, Tensor? tau, Tensor bank, Tensor labels, Tensor maxnorm, Tensor? colsum, float inv_tau, "
        "float eps, int T) -> Tensor", &bank_infonce);
  m.def("bank_large_step_loss(Tensor refer_bank, Tensor refer_idx, Tensor text, Tensor? tau, Tensor bank, Tensor labels, "
        "Tensor maxnorm, Tensor? colsum, float inv_tau, float eps, int T) -> Tensor", &bank_large_step_loss);
}
