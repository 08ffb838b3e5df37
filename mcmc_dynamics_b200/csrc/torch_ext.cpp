// PyTorch operator library over the C ABI (include/mcd_b200.h).
//
// torch is plumbing here: it owns device tensors, the current CUDA stream and (through
// torch.distributed) the NCCL communicator.  Each op forwards raw pointers and the current
// stream to the C entry point of the same name; nothing is computed in this file.
//
//   torch.ops.mcd_b200.lnprob(handle, theta[W,P] f64 cuda)         -> [W]  Runner.lnprob   (runner.py:288-306)
//   torch.ops.mcd_b200.lnlike(handle, theta)                       -> [W]  <Model>.lnlike  (constant.py:113-154, model.py:182-223, ...)
//   torch.ops.mcd_b200.lnprob_partial(handle, theta)               -> [W]  this GPU's star shard; allreduce(sum) gives lnprob
//   torch.ops.mcd_b200.lnprob_allreduce(handle, theta)             -> [W]  all shards, exchanged inside the kernel over NVLink
//   torch.ops.mcd_b200.lnlike_per_star(handle, theta[P], n_stars)  -> [N]  lnlike(no_sum=True) (model.py:620-621)
#include <ATen/ATen.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>
#include <torch/library.h>

#include "../../include/mcd_b200.h"

namespace {

mcd_handle *as_handle(int64_t h) {
    TORCH_CHECK(h != 0, "mcd_b200: null handle");
    return reinterpret_cast<mcd_handle *>(static_cast<intptr_t>(h));
}

void check_theta(const at::Tensor &theta, int64_t dims) {
    TORCH_CHECK(theta.is_cuda(), "mcd_b200: theta must be a CUDA tensor (there is no CPU path)");
    TORCH_CHECK(theta.scalar_type() == at::kDouble, "mcd_b200: theta must be float64");
    TORCH_CHECK(theta.dim() == dims, "mcd_b200: theta must have ", dims, " dimension(s)");
    TORCH_CHECK(theta.is_contiguous(), "mcd_b200: theta must be contiguous (row-major [walkers, parameters])");
}

typedef int (*ensemble_fn)(mcd_handle *, const double *, int32_t, double *, void *);

at::Tensor run(ensemble_fn fn, int64_t handle, const at::Tensor &theta) {
    check_theta(theta, 2);
    mcd_info info;
    TORCH_CHECK(mcd_get_info(as_handle(handle), &info) == 0, mcd_last_error());
    TORCH_CHECK(theta.size(1) == info.n_theta, "mcd_b200: theta has ", theta.size(1), " columns, the model has ",
                info.n_theta, " free parameters");
    // segmented handles take theta as [segments * walkers, parameters]; the C ABI wants walkers per segment
    const int64_t segments = info.n_segments > 1 ? info.n_segments : 1;
    TORCH_CHECK(theta.size(0) % segments == 0, "mcd_b200: theta rows must be a multiple of the ", segments, " segments");
    c10::cuda::CUDAGuard guard(theta.device());
    at::Tensor out = at::empty({theta.size(0)}, theta.options());
    auto stream = c10::cuda::getCurrentCUDAStream(theta.get_device());
    const int rc = fn(as_handle(handle), theta.data_ptr<double>(), (int32_t)(theta.size(0) / segments),
                      out.data_ptr<double>(), stream.stream());
    TORCH_CHECK(rc == 0, "mcd_b200: ", mcd_last_error());
    return out;
}

at::Tensor lnprob(int64_t handle, const at::Tensor &theta) { return run(mcd_lnprob_device, handle, theta); }
at::Tensor lnlike(int64_t handle, const at::Tensor &theta) { return run(mcd_lnlike_device, handle, theta); }
at::Tensor lnprob_partial(int64_t handle, const at::Tensor &theta) {
    return run(mcd_lnprob_partial_device, handle, theta);
}

at::Tensor lnprob_allreduce(int64_t handle, const at::Tensor &theta) {
    return run(mcd_lnprob_allreduce_device, handle, theta);
}

at::Tensor lnlike_per_star(int64_t handle, const at::Tensor &theta) {
    check_theta(theta, 1);
    mcd_info info;
    TORCH_CHECK(mcd_get_info(as_handle(handle), &info) == 0, mcd_last_error());
    TORCH_CHECK(theta.size(0) == info.n_theta, "mcd_b200: theta has the wrong length");
    c10::cuda::CUDAGuard guard(theta.device());
    at::Tensor out = at::empty({info.n_stars}, theta.options());
    auto stream = c10::cuda::getCurrentCUDAStream(theta.get_device());
    const int rc = mcd_lnlike_per_star_device(as_handle(handle), theta.data_ptr<double>(), out.data_ptr<double>(),
                                              stream.stream());
    TORCH_CHECK(rc == 0, "mcd_b200: ", mcd_last_error());
    return out;
}

}  // namespace

TORCH_LIBRARY(mcd_b200, m) {
    m.def("lnprob(int handle, Tensor theta) -> Tensor");
    m.def("lnlike(int handle, Tensor theta) -> Tensor");
    m.def("lnprob_partial(int handle, Tensor theta) -> Tensor");
    m.def("lnprob_allreduce(int handle, Tensor theta) -> Tensor");
    m.def("lnlike_per_star(int handle, Tensor theta) -> Tensor");
}

TORCH_LIBRARY_IMPL(mcd_b200, CUDA, m) {
    m.impl("lnprob", lnprob);
    m.impl("lnlike", lnlike);
    m.impl("lnprob_partial", lnprob_partial);
    m.impl("lnprob_allreduce", lnprob_allreduce);
    m.impl("lnlike_per_star", lnlike_per_star);
}
