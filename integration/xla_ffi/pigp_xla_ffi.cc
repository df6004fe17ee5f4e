// XLA FFI custom-call handlers over the device-pointer entry points of include/pigp.h -- the form in which
// BASELINE.json's north star asks the path to be reached from JAX ("a thin C-ABI registered as JAX FFI custom calls").
//
// NOT built by default: neither JAX nor its headers (xla/ffi/api/ffi.h) exist in this image or on the GPU box, so this
// translation unit has never been compiled here.  Where JAX >= 0.4.31 is installed:
//     make -C integration/xla_ffi JAX_FFI_INCLUDE=$(python -c "import jax.ffi; print(jax.ffi.include_dir())")
// builds libpigp_xla_ffi.so next to libpigp.so; stopro_b200/jax_ffi.py registers the targets.  The handlers only
// forward: every pigp_* entry point used here enqueues on the stream it is given and never synchronises, which is the
// FFI contract.  Plans / solvers are created from Python (set_constants) and travel as int64 attributes.
//
// Replaced reference interfaces: GPmodel.trainingK_all / mixedK_all / testK_all (GP/gp.py:287-306),
// trainingFunction_all (:213-224), d_trainingFunction_all (:412-488), predictingFunction_all (:226-256).
#include <cstdint>

#include "xla/ffi/api/ffi.h"

#include "../../include/pigp.h"

namespace ffi = xla::ffi;

static ffi::Error status(int rc) {
    if (rc == PIGP_OK) return ffi::Error::Success();
    return ffi::Error(ffi::ErrorCode::kInternal, pigp_last_error());
}

static ffi::Error AssembleImpl(cudaStream_t stream, int64_t plan, double eps, int64_t add_diag, ffi::Buffer<ffi::F64> theta,
                               ffi::ResultBuffer<ffi::F64> K) {
    const auto dims = K->dimensions();
    const int64_t ld = dims.size() == 2 ? dims[1] : 0;
    return status(pigp_assemble(reinterpret_cast<const pigp_plan*>(plan), theta.typed_data(), eps, (int)add_diag,
                                K->typed_data(), ld, PIGP_LAYOUT_FULL, stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(PigpAssemble, AssembleImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("plan")
                                  .Attr<double>("eps")
                                  .Attr<int64_t>("add_diag")
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>());

static ffi::Error NllImpl(cudaStream_t stream, int64_t solver, double eps, ffi::Buffer<ffi::F64> theta, ffi::Buffer<ffi::F64> y,
                          ffi::ResultBuffer<ffi::F64> nll, ffi::ResultBuffer<ffi::S32> info) {
    return status(pigp_nll(reinterpret_cast<pigp_solver*>(solver), theta.typed_data(), y.typed_data(), eps, nll->typed_data(),
                           info->typed_data(), stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(PigpNll, NllImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("solver")
                                  .Attr<double>("eps")
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::S32>>());

static ffi::Error NllGradImpl(cudaStream_t stream, int64_t solver, double eps, ffi::Buffer<ffi::F64> theta,
                              ffi::Buffer<ffi::F64> y, ffi::ResultBuffer<ffi::F64> nll, ffi::ResultBuffer<ffi::F64> grad,
                              ffi::ResultBuffer<ffi::S32> info) {
    return status(pigp_nll_grad(reinterpret_cast<pigp_solver*>(solver), theta.typed_data(), y.typed_data(), eps,
                                nll->typed_data(), grad->typed_data(), info->typed_data(), stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(PigpNllGrad, NllGradImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("solver")
                                  .Attr<double>("eps")
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::S32>>());

static ffi::Error PredictImpl(cudaStream_t stream, int64_t solver, int64_t mixed, int64_t test, double eps, int64_t full_cov,
                              ffi::Buffer<ffi::F64> theta, ffi::Buffer<ffi::F64> y, ffi::ResultBuffer<ffi::F64> mu,
                              ffi::ResultBuffer<ffi::F64> cov, ffi::ResultBuffer<ffi::S32> info) {
    return status(pigp_predict(reinterpret_cast<pigp_solver*>(solver), reinterpret_cast<const pigp_plan*>(mixed),
                               reinterpret_cast<const pigp_plan*>(test), theta.typed_data(), y.typed_data(), eps,
                               mu->typed_data(), cov->typed_data(), (int)full_cov, info->typed_data(), stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(PigpPredict, PredictImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("solver")
                                  .Attr<int64_t>("mixed")
                                  .Attr<int64_t>("test")
                                  .Attr<double>("eps")
                                  .Attr<int64_t>("full_cov")
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::S32>>());
