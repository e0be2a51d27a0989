// Shared device plumbing for the host classes: one bbgpu context per process/device, the
// translation of a multivariate_problem into a bbgpu objective handle, and error surfacing.
#ifndef BBOPT_HOST_DEVICE_H_
#define BBOPT_HOST_DEVICE_H_

#include <stdexcept>
#include <string>

#include "bbgpu.h"
#include "multivariate.h"

namespace bbopt {

struct GpuError: public std::runtime_error {
    explicit GpuError(const std::string &what) : std::runtime_error(what) {}
};

// Throws on a non-zero status.  When the failure is a user objective that threw inside the batched host-callback
// bridge (e.g. a Python exception, py::error_already_set), the ORIGINAL exception is rethrown, as in the reference,
// where the callback's exception propagates through optimize() (py/multivariate_py.cpp:385-388).
void gpu_check(int status);

// process-wide context on the selected device (env BBOPT_DEVICE or LOCAL_RANK, default 0).
// Throws when no CUDA device is usable: there is no CPU fallback.
bbgpu_ctx *device_context();

// RAII objective handle built from a problem: registry / user CUDA objective when `_dev` is set,
// otherwise the batched host-callback bridge around `_f`.
class ObjectiveHandle {
public:
    ObjectiveHandle() {}
    ~ObjectiveHandle() { reset(); }
    ObjectiveHandle(const ObjectiveHandle&) = delete;
    ObjectiveHandle &operator=(const ObjectiveHandle&) = delete;
    void bind(const multivariate_problem &p);
    void reset();
    bbgpu_objective *get() const { return _h; }
    // single evaluation of a host point (the restart drivers' "+1 evaluation")
    double evalOne(const double *x);
private:
    bbgpu_objective *_h = nullptr;
    multivariate _f; // kept alive for the host bridge
    int _n = 0;
    static int bridge(const double *hX, int64_t ld, int n, int count, double *hf, void *user);
};

} // namespace bbopt

#endif
