// placeholder until de.cu lands (keeps every symbol of include/bbgpu.h exported)
#include "common.cuh"
extern "C" {
int bbgpu_de_create(bbgpu_ctx *, const bbgpu_de_config *, const double *, const double *, bbgpu_de **) { return bb::fail(BBGPU_ERR_ARG, "DE engine not built"); }
int bbgpu_de_destroy(bbgpu_de *) { return BBGPU_OK; }
int bbgpu_de_init(bbgpu_de *, bbgpu_objective *, const double *) { return bb::fail(BBGPU_ERR_ARG, "DE engine not built"); }
int bbgpu_de_generation(bbgpu_de *, bbgpu_objective *, const bbgpu_de_draws *, bbgpu_de_report *) { return bb::fail(BBGPU_ERR_ARG, "DE engine not built"); }
int bbgpu_de_get_swarm(bbgpu_de *, double *, double *) { return bb::fail(BBGPU_ERR_ARG, "DE engine not built"); }
int bbgpu_de_get_archive(bbgpu_de *, double *, int *) { return bb::fail(BBGPU_ERR_ARG, "DE engine not built"); }
int bbgpu_de_get_memory(bbgpu_de *, double *, double *) { return bb::fail(BBGPU_ERR_ARG, "DE engine not built"); }
int bbgpu_de_get_trace(bbgpu_de *, int32_t *, double *, double *, double *) { return bb::fail(BBGPU_ERR_ARG, "DE engine not built"); }
}
