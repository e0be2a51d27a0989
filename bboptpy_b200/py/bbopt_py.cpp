// pybind11 module `_bboptpy` of the accelerated path.
//
// Exposes the reference's Python surface for the classes on the hot path with the same class
// names, constructor keywords and defaults (/root/reference/py/multivariate_py.cpp:99-171) and
// the same generic methods optimize / initialize / iterate / solution (:357-420).  `f` may be
//   * a registry objective (bboptpy_b200.Rosenbrock, ...) or a user CUDA objective -> evaluated on
//     the GPU, the whole population per launch;
//   * any Python callable f(x: ndarray) -> float, exactly as in the reference -> driven through the
//     batched host-callback bridge (one device->host copy per generation).
#include <pybind11/functional.h>
#include <pybind11/iostream.h>
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <execinfo.h>
#include <signal.h>
#include <unistd.h>

#include "cma.h"
#include "de.h"
#include "sansde.h"
#include "ssde.h"
#include "xnes.h"
#include "restart.h"
#include "batch.h"

namespace py = pybind11;
using namespace pybind11::literals;

namespace {

multivariate_problem make_problem(py::object &py_f, py::array_t<double> &py_lower, py::array_t<double> &py_upper) {
    const int n = (int) py_lower.size();
    double *lower = static_cast<double*>(py_lower.request().ptr);
    double *upper = static_cast<double*>(py_upper.request().ptr);
    if (py::isinstance<device_objective>(py_f)) {
        return multivariate_problem(py_f.cast<device_objective>(), n, lower, upper);
    }
    if (!PyCallable_Check(py_f.ptr())) throw py::type_error("f must be a registry/CUDA Objective or a callable f(x) -> float");
    py::function fn = py::reinterpret_borrow<py::function>(py_f);
    multivariate f = [fn, n](const double *x) -> double {
        const auto py_x = py::array_t<double>(n, x);
        return fn(py_x).cast<double>();
    };
    return multivariate_problem(f, n, lower, upper);
}

py::array_t<double> as_array(const std::vector<double> &v) {
    return py::array_t<double>(v.size(), v.data());
}

} // namespace

namespace {
// BBOPT_BACKTRACE=1: print a native backtrace on SIGSEGV (debug aid; off by default)
void segv_backtrace(int sig) {
    void *frames[64];
    const int n = backtrace(frames, 64);
    backtrace_symbols_fd(frames, n, STDERR_FILENO);
    signal(sig, SIG_DFL);
    raise(sig);
}
} // namespace

PYBIND11_MODULE(_bboptpy, m) {
    m.doc() = "B200-native population engine behind bboptpy's optimizer API";
    if (std::getenv("BBOPT_BACKTRACE")) signal(SIGSEGV, segv_backtrace);
    m.def("seed", [](unsigned s) { bbopt::Random::seed(s); }, "s"_a,
            "Seed the host engine and the device Philox streams (the reference exposes no seed).");
    m.def("uniform", [](double a, double b) { return bbopt::Random::get(a, b); }, "a"_a, "b"_a,
            "One uniform_real draw from the host engine (the restart drivers' stream).");

    py::class_<device_objective>(m, "Objective")
        .def_static("builtin", [](int id) { device_objective d; d._builtin = id; return d; }, "id"_a)
        .def_static("from_pointer", [](std::uintptr_t fn, std::uintptr_t user) {
            device_objective d;
            d._fn = reinterpret_cast<bbgpu_device_objective_fn>(fn);
            d._user = reinterpret_cast<void*>(user);
            return d;
        }, "fn"_a, "user"_a = 0,
        "User CUDA objective: address of `int fn(const double* d_X, int64 ld, int n, int count, double* d_f, void* stream, void* user)`")
        .def_property_readonly("id", [](const device_objective &d) { return d._builtin; });
    auto reg = [&](const char *name, int id) { device_objective d; d._builtin = id; m.attr(name) = d; };
    reg("Sphere", BBGPU_OBJ_SPHERE);
    reg("Ellipsoid", BBGPU_OBJ_ELLIPSOID);
    reg("Rosenbrock", BBGPU_OBJ_ROSENBROCK);
    reg("Rastrigin", BBGPU_OBJ_RASTRIGIN);
    reg("Ackley", BBGPU_OBJ_ACKLEY);

    py::class_<multivariate_solution> solution(m, "MultivariateSolution");
    solution.def("__str__", &multivariate_solution::toString);
    solution.def_property_readonly("x", [](multivariate_solution &self) { return as_array(self._sol); });
    solution.def_property_readonly("converged", [](multivariate_solution &self) { return self._converged; });
    solution.def_property_readonly("n_evals", [](multivariate_solution &self) { return self._fev; });

    py::class_<MultivariateOptimizer> solver(m, "MultivariateSearch");
    solver.def("optimize", [](MultivariateOptimizer &self, py::object py_f, py::array_t<double> py_lower,
            py::array_t<double> py_upper, py::array_t<double> py_guess) {
        multivariate_problem prob = make_problem(py_f, py_lower, py_upper);
        double *guess = static_cast<double*>(py_guess.request().ptr);
        return self.optimize(prob, guess);
    }, "f"_a, "lower"_a, "upper"_a, "guess"_a, py::call_guard<py::scoped_ostream_redirect, py::scoped_estream_redirect>());
    solver.def("initialize", [](MultivariateOptimizer &self, py::object py_f, py::array_t<double> py_lower,
            py::array_t<double> py_upper, py::array_t<double> py_guess) {
        multivariate_problem prob = make_problem(py_f, py_lower, py_upper);
        double *guess = static_cast<double*>(py_guess.request().ptr);
        return self.init(prob, guess);
    }, "f"_a, "lower"_a, "upper"_a, "guess"_a, py::call_guard<py::scoped_ostream_redirect, py::scoped_estream_redirect>());
    solver.def("iterate", &MultivariateOptimizer::iterate);
    solver.def("solution", &MultivariateOptimizer::solution);

    // one rank of an NCCL communicator (bbgpu_comm_*): candidate-sharded runs and best-of reduces
    struct PyComm {
        bbgpu_comm *h = nullptr;
        PyComm(int world, int rank, py::bytes id) {
            std::string raw = id;
            if ((int) raw.size() != BBGPU_COMM_ID_BYTES) throw std::invalid_argument("unique id must be 128 bytes");
            bbopt::gpu_check(bbgpu_comm_create(bbopt::device_context(), world, rank, raw.data(), &h));
        }
        ~PyComm() { if (h) bbgpu_comm_destroy(h); }
    };
    py::class_<PyComm>(m, "Communicator")
        .def(py::init<int, int, py::bytes>(), "world"_a, "rank"_a, "unique_id"_a)
        .def_static("unique_id", []() {
            char buf[BBGPU_COMM_ID_BYTES];
            bbopt::gpu_check(bbgpu_comm_unique_id(buf));
            return py::bytes(buf, BBGPU_COMM_ID_BYTES);
        }, "rank 0 creates the id and hands it to the other ranks (e.g. torch.distributed.broadcast_object_list)")
        .def_property_readonly("rank", [](PyComm &c) { return bbgpu_comm_rank(c.h); })
        .def_property_readonly("world", [](PyComm &c) { return bbgpu_comm_world(c.h); });

    py::class_<BaseCmaes, MultivariateOptimizer>(m, "BaseCMAES")
        .def("set_communicator", [](BaseCmaes &self, PyComm &c, uint64_t stream_key) {
            self.setCommunicator(c.h);
            self.setStreamKey(stream_key);
        }, "comm"_a, "stream_key"_a = 0, py::keep_alive<1, 2>(),
             "Shard the candidates of this run over the ranks of `comm` (call on every rank, before optimize/initialize)")
        .def("set_stream_key", &BaseCmaes::setStreamKey, "key"_a, "Philox key of the free-running draws (default: a fresh key per object)")
        .def("state", [](BaseCmaes &self, const std::string &name) { return as_array(self.state(name)); }, "name"_a)
        .def("checkpoint", [](BaseCmaes &self) { return py::bytes(self.checkpoint()); },
             "Versioned state blob of the run (host scalars + device state); see INTEGRATION.md")
        .def("restore", [](BaseCmaes &self, const py::bytes &blob) { self.restore(std::string(blob)); }, "blob"_a,
             "Resume from checkpoint(): call after initialize(f, lower, upper, guess) on an optimizer of the same class")
        .def("set_draw_replay", &BaseCmaes::setDrawReplay, "seed"_a,
             "Draw replay: consume the reference's N(0,1) stream (std::mt19937(seed) + normal_distribution) in its own draw order")
        .def("set_draw_replay_global", &BaseCmaes::setDrawReplayGlobal)
        .def("burn_uniforms", &BaseCmaes::burnUniforms, "count"_a)
        .def("converged", &BaseCmaes::converged)
        .def_property_readonly("sigma", &BaseCmaes::sigma)
        .def_property_readonly("n_evals", &BaseCmaes::evaluations)
        .def_property_readonly("n_iters", &BaseCmaes::iterations)
        .def_property_readonly("popsize", &BaseCmaes::lambda);

    py::class_<Cmaes, BaseCmaes>(m, "CMAES")
        .def(py::init<int, double, int, double, bool, double>(), "mfev"_a, "tol"_a, "np"_a, "sigma0"_a = 2.,
             "bound"_a = false, "eigenrate"_a = 0.25)
        .def("set_eigen_replay", [](Cmaes &self, py::object fn) {
            if (fn.is_none()) { self.setEigenReplay(nullptr); return; }
            self.setEigenReplay([fn](int fev, int n, std::vector<double> &B, std::vector<double> &D) {
                py::object r = fn(fev, n);
                if (r.is_none()) return false;
                py::tuple t = r.cast<py::tuple>();
                auto b = t[0].cast<py::array_t<double, py::array::c_style | py::array::forcecast>>();
                auto d = t[1].cast<py::array_t<double, py::array::c_style | py::array::forcecast>>();
                B.assign(b.data(), b.data() + b.size());
                D.assign(d.data(), d.data() + d.size());
                return true;
            });
        }, "fn"_a, "Eigenbasis replay: fn(fev, n) -> None or (B, D) imposed at each decomposition event")
        .def_property_readonly("flag", &Cmaes::flag)
        .def_property_readonly("n_decompositions", &Cmaes::decompositions);

    py::class_<ActiveCmaes, Cmaes>(m, "ActiveCMAES")
        .def(py::init<int, double, int, double, bool, double, double>(), "mfev"_a, "tol"_a, "np"_a, "sigma0"_a = 2.,
             "bound"_a = false, "alphacov"_a = 2., "eigenrate"_a = 0.25);

    py::class_<SepCmaes, BaseCmaes>(m, "SepCMAES")
        .def(py::init<int, double, int, double, bool, bool>(), "mfev"_a, "tol"_a, "np"_a, "sigma0"_a = 2.,
             "bound"_a = false, "adjustlr"_a = true)
        .def_property_readonly("flag", &SepCmaes::flag);

    // py/multivariate_py.cpp:117-121
    py::class_<CholeskyCmaes, BaseCmaes>(m, "CholeskyCMAES")
        .def(py::init<int, double, double, int, double, bool>(), "mfev"_a, "tol"_a, "stol"_a, "np"_a, "sigma0"_a = 2.,
             "bound"_a = false);

    // py/multivariate_py.cpp:123-128
    py::class_<LmCmaes, BaseCmaes>(m, "LmCMAES")
        .def(py::init<int, double, int, int, double, bool, bool, bool>(), "mfev"_a, "tol"_a, "np"_a, "memory"_a = 0, "sigma0"_a = 2.,
             "bound"_a = false, "rademacher"_a = true, "usenew"_a = true)
        .def_property_readonly("memory_length", &LmCmaes::memoryLength);

    py::class_<IPopCmaes, MultivariateOptimizer>(m, "IPopCMAES")
        .def(py::init<BaseCmaes*, int, bool, double, bool, double, bool>(), "base"_a, "mfev"_a, "print"_a = false,
             "sigma0"_a = 2., "nipop"_a = true, "ksigmadec"_a = 1.6, "boundlambda"_a = true, py::keep_alive<1, 2>())
        .def_property_readonly("n_evals", &IPopCmaes::evaluations);

    py::class_<BiPopCmaes, MultivariateOptimizer>(m, "BiPopCMAES")
        .def(py::init<BaseCmaes*, int, bool, double, int, bool, double, double>(), "base"_a, "mfev"_a, "print"_a = false,
             "sigma0"_a = 2., "maxlargeruns"_a = 9, "nbipop"_a = true, "ksigmadec"_a = 1.6, "kbudget"_a = 2.,
             py::keep_alive<1, 2>())
        .def_property_readonly("n_evals", &BiPopCmaes::evaluations)
        .def_property_readonly("large_restarts", &BiPopCmaes::largeRestarts)
        .def_property_readonly("small_restarts", &BiPopCmaes::smallRestarts);

    // batches of independent IPOP / BIPOP runs on the batched small-n engine (an addition: the reference's unit of
    // work is one optimizer object).  Keywords follow IPopCMAES / BiPopCMAES; `tol` and `eigenrate` are the base
    // CMAES's own (its np / sigma0 / mfev are set by the restart schedule, cma/bipop_cmaes.cpp:80,215,251).
    auto batch_optimize = [](RestartBatch &self, py::object py_f, py::array_t<double> py_lower, py::array_t<double> py_upper,
            py::array_t<double, py::array::c_style | py::array::forcecast> py_guesses) {
        multivariate_problem prob = make_problem(py_f, py_lower, py_upper);
        if (py_guesses.size() != (py::ssize_t) self.totalRuns() * prob._n)
            throw std::invalid_argument("guesses must have shape (runs, n)");
        return self.optimize(prob, py_guesses.data());
    };
    py::class_<restart_batch_stats>(m, "RestartBatchStats")
        .def_readonly("base_runs", &restart_batch_stats::base_runs)
        .def_readonly("generations", &restart_batch_stats::generations)
        .def_readonly("evaluations", &restart_batch_stats::evaluations)
        .def_readonly("decompositions", &restart_batch_stats::decompositions)
        .def_readonly("launches", &restart_batch_stats::launches)
        .def_readonly("device_ms", &restart_batch_stats::device_ms);
    py::class_<RestartBatch>(m, "RestartBatch")
        .def("optimize", batch_optimize, "f"_a, "lower"_a, "upper"_a, "guesses"_a,
             "Run every run of the batch (this rank's share of it) to completion; returns their MultivariateSolutions")
        .def("set_communicator", [](RestartBatch &self, PyComm &c) { self.setCommunicator(c.h); }, "comm"_a, py::keep_alive<1, 2>(),
             "Shard the runs over the ranks of `comm`: this process runs the runs r with r % world == rank")
        .def("set_quantum", &RestartBatch::setQuantum, "evals"_a, "Evaluations a run may spend per launch")
        .def("set_timing", &RestartBatch::setTiming, "on"_a)
        .def("set_streams", &RestartBatch::setStreams, "streams"_a, "Concurrent streams the runs are dealt to (0 = automatic)")
        .def_property_readonly("run_ids", [](RestartBatch &self) { return self.runIds(); })
        .def_property_readonly("best_values", [](RestartBatch &self) { return as_array(self.bestValues()); })
        .def_property_readonly("stats", [](RestartBatch &self) { return self.stats(); })
        .def("best_of", [](RestartBatch &self) {
            std::vector<double> x;
            int owner = -1;
            const double f = self.bestOf(x, &owner);
            return py::make_tuple(f, owner, as_array(x));
        }, "Best run of the whole batch over all ranks (one NCCL best-of reduce): (f, global run id, x)");
    struct BiPopBatch: public RestartBatch { using RestartBatch::RestartBatch; };
    struct IPopBatch: public RestartBatch { using RestartBatch::RestartBatch; };
    py::class_<BiPopBatch, RestartBatch>(m, "BiPopCMAESBatch")
        .def(py::init([](int runs, int mfev, double tol, double sigma0, int maxlargeruns, bool nbipop, double ksigmadec,
                double kbudget, double eigenrate) {
            auto *b = new BiPopBatch(RestartBatch::BIPOP, runs, mfev, tol, sigma0, eigenrate);
            b->configureBiPop(maxlargeruns, nbipop, ksigmadec, kbudget);
            return b;
        }), "runs"_a, "mfev"_a, "tol"_a, "sigma0"_a = 2., "maxlargeruns"_a = 9, "nbipop"_a = true, "ksigmadec"_a = 1.6,
            "kbudget"_a = 2., "eigenrate"_a = 0.25);
    py::class_<IPopBatch, RestartBatch>(m, "IPopCMAESBatch")
        .def(py::init([](int runs, int mfev, double tol, double sigma0, bool nipop, double ksigmadec, bool boundlambda,
                double eigenrate) {
            auto *b = new IPopBatch(RestartBatch::IPOP, runs, mfev, tol, sigma0, eigenrate);
            b->configureIPop(nipop, ksigmadec, boundlambda);
            return b;
        }), "runs"_a, "mfev"_a, "tol"_a, "sigma0"_a = 2., "nipop"_a = true, "ksigmadec"_a = 1.6, "boundlambda"_a = true,
            "eigenrate"_a = 0.25);

    // py/multivariate_py.cpp:153-157
    py::class_<xNES, MultivariateOptimizer>(m, "xNES")
        .def(py::init<int, double, double, double>(), "mfev"_a, "tol"_a, "a0"_a = 1., "etamu"_a = 1.)
        .def("set_draw_replay", &xNES::setDrawReplay, "seed"_a)
        .def("state", [](xNES &self, const std::string &name) { return as_array(self.state(name)); }, "name"_a)
        .def("converged", &xNES::isConverged)
        .def_property_readonly("sigma", &xNES::sigma)
        .def_property_readonly("gsigma", &xNES::gsigma)
        .def_property_readonly("n_evals", &xNES::evaluations)
        .def_property_readonly("np", &xNES::population);

    // py/multivariate_py.cpp:173-178
    py::class_<SaNSDESearch, MultivariateOptimizer>(m, "SANSDE")
        .def(py::init<int, int, double, bool, int, int, int>(), "mfev"_a, "np"_a, "tol"_a, "repaircr"_a = true, "crref"_a = 5,
             "pupdate"_a = 50, "crupdate"_a = 25)
        .def("swarm", [](SaNSDESearch &self) {
            std::vector<double> x = self.swarm();
            const int np = self.population();
            py::array_t<double> a({ (py::ssize_t) np, (py::ssize_t) (np ? x.size() / np : 0) });
            std::copy(x.begin(), x.end(), a.mutable_data());
            return a;
        })
        .def("fitness", [](SaNSDESearch &self) { return as_array(self.fitness()); })
        .def("crossover_rates", [](SaNSDESearch &self) { return as_array(self.crossoverRates()); })
        .def("set_stream_key", &SaNSDESearch::setStreamKey, "key"_a)
        .def("set_initial_population", [](SaNSDESearch &self, py::array_t<double, py::array::c_style | py::array::forcecast> x) {
            self.setInitialPopulation(std::vector<double>(x.data(), x.data() + x.size()));
        }, "x"_a)
        .def("set_draw_replay", [](SaNSDESearch &self, py::object fn) {
            if (fn.is_none()) { self.setDrawReplay(nullptr); return; }
            self.setDrawReplay([fn](int gen, sansde_generation_draws &d) {
                py::object r = fn(gen);
                if (r.is_none()) return false;
                py::dict t = r.cast<py::dict>();
                auto dv = [&](const char *k, std::vector<double> &out) {
                    if (!t.contains(k)) return;
                    auto a = t[k].cast<py::array_t<double, py::array::c_style | py::array::forcecast>>();
                    out.assign(a.data(), a.data() + a.size());
                };
                auto iv = [&](const char *k, std::vector<int32_t> &out) {
                    if (!t.contains(k)) return;
                    auto a = t[k].cast<py::array_t<double, py::array::c_style | py::array::forcecast>>();
                    out.resize(a.size());
                    for (py::ssize_t i = 0; i < a.size(); i++) out[i] = (int32_t) a.data()[i];
                };
                dv("zcr", d.zcr); dv("F", d.F); dv("u", d.u);
                iv("ifstrat", d.ifstrat); iv("r1", d.r1); iv("r2", d.r2); iv("r3", d.r3); iv("istrat", d.istrat); iv("jrand", d.jrand);
                return true;
            });
        }, "fn"_a)
        .def("converged", &SaNSDESearch::isConverged)
        .def("sums", [](SaNSDESearch &self) {
            const bbgpu_sde_accum &a = self.sums();
            py::dict d;
            d["pns0"] = a.pns[0]; d["pns1"] = a.pns[1]; d["pnf0"] = a.pnf[0]; d["pnf1"] = a.pnf[1];
            d["crrec"] = a.crrec; d["crdeltaf"] = a.crdeltaf; d["fpns0"] = a.fpns[0]; d["fpns1"] = a.fpns[1];
            d["fpnf0"] = a.fpnf[0]; d["fpnf1"] = a.fpnf[1]; d["nsucc"] = a.nsucc;
            return d;
        })
        .def_property_readonly("p", &SaNSDESearch::p)
        .def_property_readonly("fp", &SaNSDESearch::fp)
        .def_property_readonly("crm", &SaNSDESearch::crm)
        .def_property_readonly("rounds", &SaNSDESearch::rounds)
        .def_property_readonly("n_evals", &SaNSDESearch::evaluations)
        .def_property_readonly("popsize", &SaNSDESearch::population);

    // py/multivariate_py.cpp:180-186 of the reference (note its Python default usede = False)
    py::class_<SSDESearch, MultivariateOptimizer>(m, "SSDE")
        .def(py::init<int, int, double, int, int, double, int, bool, bool>(), "mfev"_a, "npinit"_a, "tol"_a, "patience"_a = 1000,
             "npmin"_a = 4, "ptop"_a = 0.11, "h"_a = 100, "usede"_a = false, "repaircr"_a = true)
        .def("swarm", [](SSDESearch &self) {
            std::vector<double> x = self.swarm();
            const int np = self.population();
            py::array_t<double> a({ (py::ssize_t) np, (py::ssize_t) (np ? x.size() / np : 0) });
            std::copy(x.begin(), x.end(), a.mutable_data());
            return a;
        })
        .def("fitness", [](SSDESearch &self) { return as_array(self.fitness()); })
        .def("memory", [](SSDESearch &self, int which) { return as_array(self.memory(which)); }, "which"_a)
        .def("trace", [](SSDESearch &self) {
            std::vector<int32_t> succ;
            std::vector<double> ft;
            self.trace(succ, ft);
            py::array_t<int32_t> s((py::ssize_t) succ.size());
            std::copy(succ.begin(), succ.end(), s.mutable_data());
            return py::make_tuple(s, as_array(ft));
        })
        .def("set_stream_key", &SSDESearch::setStreamKey, "key"_a)
        .def("set_initial_population", [](SSDESearch &self, py::array_t<double, py::array::c_style | py::array::forcecast> x) {
            self.setInitialPopulation(std::vector<double>(x.data(), x.data() + x.size()));
        }, "x"_a)
        .def("set_draw_replay", [](SSDESearch &self, py::object fn) {
            if (fn.is_none()) { self.setDrawReplay(nullptr); return; }
            self.setDrawReplay([fn](int gen, ssde_generation_draws &d) {
                py::object r = fn(gen);
                if (r.is_none()) return false;
                py::dict t = r.cast<py::dict>();
                auto dv = [&](const char *k, std::vector<double> &out) {
                    if (!t.contains(k)) return;
                    auto a = t[k].cast<py::array_t<double, py::array::c_style | py::array::forcecast>>();
                    out.assign(a.data(), a.data() + a.size());
                };
                auto iv = [&](const char *k, std::vector<int32_t> &out) {
                    if (!t.contains(k)) return;
                    auto a = t[k].cast<py::array_t<double, py::array::c_style | py::array::forcecast>>();
                    out.resize(a.size());
                    for (py::ssize_t i = 0; i < a.size(); i++) out[i] = (int32_t) a.data()[i];
                };
                iv("iwork", d.iwork); iv("iL", d.iL); iv("p", d.p); iv("q", d.q); iv("r", d.r); iv("pbest", d.pbest); iv("j0", d.j0);
                dv("zrank", d.zrank); dv("ub", d.ub); dv("cauchy", d.cauchy); dv("zcr", d.zcr); dv("ucr", d.ucr); dv("redraw", d.redraw);
                return true;
            });
        }, "fn"_a)
        .def("converged", &SSDESearch::isConverged)
        .def_property_readonly("rounds", &SSDESearch::rounds)
        .def_property_readonly("accepted", &SSDESearch::accepted)
        .def_property_readonly("convcount", &SSDESearch::convcount)
        .def_property_readonly("k", &SSDESearch::k)
        .def_property_readonly("kCR", &SSDESearch::kCR)
        .def_property_readonly("n_evals", &SSDESearch::evaluations)
        .def_property_readonly("popsize", &SSDESearch::population);

    py::class_<DeSearchBase, MultivariateOptimizer>(m, "BaseDE")
        .def("swarm", [](DeSearchBase &self) {
            std::vector<double> x = self.swarm();
            const int np = self.population();
            py::array_t<double> a({ (py::ssize_t) np, (py::ssize_t) (np ? x.size() / np : 0) });
            std::copy(x.begin(), x.end(), a.mutable_data());
            return a;
        })
        .def("fitness", [](DeSearchBase &self) { return as_array(self.fitness()); })
        .def("set_stream_key", &DeSearchBase::setStreamKey, "key"_a,
             "Fix the Philox key of the free-running draws (default: a fresh key per initialize()/optimize(), so repeated calls are independent runs)")
        .def("set_initial_population", [](DeSearchBase &self, py::array_t<double, py::array::c_style | py::array::forcecast> x) {
            self.setInitialPopulation(std::vector<double>(x.data(), x.data() + x.size()));
        }, "x"_a)
        .def("set_draw_replay", [](DeSearchBase &self, py::object fn) {
            if (fn.is_none()) { self.setDrawReplay(nullptr); return; }
            self.setDrawReplay([fn](int gen, de_generation_draws &d) {
                py::object r = fn(gen);
                if (r.is_none()) return false;
                py::dict t = r.cast<py::dict>();
                auto dv = [&](const char *k, std::vector<double> &out) {
                    if (!t.contains(k)) return;
                    auto a = t[k].cast<py::array_t<double, py::array::c_style | py::array::forcecast>>();
                    out.assign(a.data(), a.data() + a.size());
                };
                auto iv = [&](const char *k, std::vector<int32_t> &out) {
                    if (!t.contains(k)) return;
                    auto a = t[k].cast<py::array_t<double, py::array::c_style | py::array::forcecast>>();
                    out.resize(a.size());
                    for (py::ssize_t i = 0; i < a.size(); i++) out[i] = (int32_t) a.data()[i];
                };
                dv("zcr", d.zcr); dv("cauchy", d.cauchy); dv("u", d.u); dv("pi", d.pi);
                iv("ibest", d.ibest); iv("r1", d.r1); iv("r2", d.r2); iv("jrand", d.jrand); iv("slot", d.slot);
                iv("ri", d.ri); iv("trim", d.trim);
                return true;
            });
        }, "fn"_a, "Draw replay: fn(generation) -> None or the oracle recorder's tagged-draw dict")
        .def_property_readonly("n_evals", &DeSearchBase::evaluations)
        .def_property_readonly("popsize", &DeSearchBase::population)
        .def_property_readonly("rounds", [](DeSearchBase &self) { return self.report().rounds; })
        .def_property_readonly("mucr", [](DeSearchBase &self) { return self.report().mucr; })
        .def_property_readonly("muf", [](DeSearchBase &self) { return self.report().muf; });

    py::class_<JadeSearch, DeSearchBase>(m, "JADE")
        .def(py::init<int, int, double, bool, bool, double, double, double>(), "mfev"_a, "np"_a, "tol"_a,
             "archive"_a = true, "repaircr"_a = true, "pelite"_a = 0.05, "cdamp"_a = 0.1, "sigma"_a = 0.07);

    py::class_<ShadeSearch, DeSearchBase>(m, "SHADE")
        .def(py::init<int, int, double, bool, bool, int, int>(), "mfev"_a, "npinit"_a, "tol"_a, "archive"_a = true,
             "repaircr"_a = true, "h"_a = 100, "npmin"_a = 4);
}
