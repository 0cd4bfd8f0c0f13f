/*
 * ref_pso_shim.cpp — thin harness around the UNMODIFIED reference PSO (TMVS/pso/psosolver.cpp, particle.cpp),
 * compiled in place from /root/reference by oracle/Makefile into oracle/_ref/libpso_ref.so.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/pmvs_oracle.cpp header). No reference source is copied into the repo.
 *
 * The reference seeds with srand(time(NULL)+tid) and draws rand() (psosolver.cpp:60-68), which is not
 * reproducible. This shim interposes srand()/rand() (the library is linked -Bsymbolic so the reference's calls bind
 * here) with the repo's counter-based stream: draw k of a solver returns rand31(key, k), RAND_MAX = 2^31-1 (glibc).
 */
#include <atomic>
#include <cstdint>
#include <cstring>
#include <omp.h>
#include "psosolver.h"

namespace {
const uint64_t GOLD = 0x9E3779B97F4A7C15ULL;
inline uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
struct Stream { uint64_t key; std::atomic<uint64_t> ctr; };
thread_local Stream tlsStream;
Stream sharedStream;
bool useShared = false;

typedef double (*FitFn)(const double *pos, void *obj);
struct Bundle { FitFn fn; void *obj; };
double adapter(const PAIS::Particle &p, void *obj) {
    Bundle *b = (Bundle *)obj;
    return b->fn(p.pos, b->obj);
}
}

extern "C" {

int rand(void) {
    Stream &s = useShared ? sharedStream : tlsStream;
    uint64_t c = s.ctr.fetch_add(1, std::memory_order_relaxed);
    return (int)(mix64(s.key + GOLD * (c + 1)) >> 33);
}
void srand(unsigned int) { /* the stream is keyed by the caller, psosolver.cpp:60-64 is neutralised */ }

/* Runs `new PsoSolver(3, L, U, fn, obj, maxIter, P)` -> setParticle(init) -> run(true) exactly as
 * Patch::psoOptimization does (TMVS/mvs/patch.cpp:190-213). nThreads > 1 enables the reference's own OpenMP loops
 * (draw order then depends on thread scheduling, as in the reference). */
int ref_pso_solve(const double *L, const double *U, FitFn fn, void *obj, int maxIter, int P, const double *init, uint64_t key,
                  double *gbest, double *gbestFitness, int *iterations, int nThreads) {
    useShared = nThreads > 1;
    Stream &s = useShared ? sharedStream : tlsStream;
    s.key = key;
    s.ctr.store(0);
    omp_set_num_threads(nThreads > 1 ? nThreads : 1);
    Bundle b = {fn, obj};
    PAIS::PsoSolver *solver = new PAIS::PsoSolver(3, L, U, adapter, &b, maxIter, P);
    if (init) solver->setParticle(init);
    solver->run(true);
    const double *g = solver->getGbest();
    memcpy(gbest, g, 3 * sizeof(double));
    *gbestFitness = solver->getGbestFitness();
    *iterations = solver->getIteration();
    delete solver;
    return 0;
}

int ref_pso_solve_basic(const double *L, const double *U, FitFn fn, void *obj, int maxIter, int P, const double *init, uint64_t key,
                        int glnpso, double *gbest, double *gbestFitness, int *iterations) {
    useShared = false;
    tlsStream.key = key;
    tlsStream.ctr.store(0);
    omp_set_num_threads(1);
    Bundle b = {fn, obj};
    PAIS::PsoSolver *solver = new PAIS::PsoSolver(3, L, U, adapter, &b, maxIter, P);
    if (init) solver->setParticle(init);
    solver->run(glnpso != 0);
    memcpy(gbest, solver->getGbest(), 3 * sizeof(double));
    *gbestFitness = solver->getGbestFitness();
    *iterations = solver->getIteration();
    delete solver;
    return 0;
}

void *ref_pso_solve_ptr(void) { return (void *)ref_pso_solve; }
}
