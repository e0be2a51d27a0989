#include "device.h"

#include <cstdlib>
#include <exception>
#include <mutex>

namespace bbopt {

// the exception a host objective threw inside the C call that is returning an error on this thread
static thread_local std::exception_ptr pending_objective_exception;

void gpu_check(int status) {
    if (status == BBGPU_OK) return;
    if (pending_objective_exception) {
        std::exception_ptr e = pending_objective_exception;
        pending_objective_exception = nullptr;
        std::rethrow_exception(e);
    }
    throw GpuError(std::string("bbgpu: ") + bbgpu_last_error());
}

bbgpu_ctx *device_context() {
    static bbgpu_ctx *ctx = nullptr;
    static std::mutex m;
    std::lock_guard<std::mutex> lock(m);
    if (!ctx) {
        int dev = 0;
        if (const char *e = std::getenv("BBOPT_DEVICE")) dev = std::atoi(e);
        else if (const char *r = std::getenv("LOCAL_RANK")) dev = std::atoi(r);
        gpu_check(bbgpu_ctx_create(dev, nullptr, &ctx));
    }
    return ctx;
}

int ObjectiveHandle::bridge(const double *hX, int64_t ld, int n, int count, double *hf, void *user) {
    (void) n;
    ObjectiveHandle *self = static_cast<ObjectiveHandle*>(user);
    try {
        for (int k = 0; k < count; k++) hf[k] = self->_f(hX + (int64_t) k * ld);
    } catch (...) {
        pending_objective_exception = std::current_exception(); // rethrown by gpu_check once the C call has returned
        return 1;
    }
    return 0;
}

void ObjectiveHandle::bind(const multivariate_problem &p) {
    reset();
    _n = p._n;
    bbgpu_ctx *ctx = device_context();
    if (p._dev._builtin >= 0) {
        gpu_check(bbgpu_objective_builtin(ctx, p._dev._builtin, p._n, &_h));
    } else if (p._dev._fn != nullptr) {
        gpu_check(bbgpu_objective_device(ctx, p._dev._fn, p._dev._user, p._n, &_h));
    } else if (p._f) {
        _f = p._f;
        gpu_check(bbgpu_objective_host(ctx, &ObjectiveHandle::bridge, this, p._n, &_h));
    } else {
        throw GpuError("multivariate_problem has neither a device objective nor a host function");
    }
}

void ObjectiveHandle::reset() {
    if (_h) bbgpu_objective_destroy(_h);
    _h = nullptr;
}

double ObjectiveHandle::evalOne(const double *x) {
    double f = 0.0;
    gpu_check(bbgpu_eval_host(device_context(), _h, x, 1, &f));
    return f;
}

} // namespace bbopt
