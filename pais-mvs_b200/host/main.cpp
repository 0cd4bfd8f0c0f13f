/*
 * tmvs — command-line driver: `tmvs -r <file.nvm|.nvm2|.mvs>` is the reference's `TMVS.exe -r` (TMVS/TMVS.cpp:76-122,
 * :174-203): compiled defaults -> config.txt -> load -> config.txt again -> init.mvs -> seed refinement -> seed.mvs ->
 * expansion -> exp.mvs / exp.ply / exp.psr, total time printed as "time1". `tmvs -f <file.mvs>` is `TMVS.exe -f`
 * (TMVS.cpp:124-170). The viewer commands (-v / -a) are outside this repository's scope.
 *
 * Extra switches (not in the reference): --config FILE, --out-dir DIR, --image-dir DIR, --round K (parents per expansion
 * round, default 1024), --device D, --gpus N (shard every batch over N GPUs), --seed S (run seed of the counter-based PSO
 * RNG), --merge-slots / --slot-passes (one GPU pass per round over all camera slots, or one per slot: the reference's visiting order),
 * --no-pipeline (merged mode: one round at a time instead of two in flight — commit k / generate k+2 overlapping GPU round k+1),
 * --autosave-seconds T (spacing of auto_save.mvs checkpoints, default 5), --no-expand, -V (verbose),
 * --convert IN OUT.{mvs,ply,psr} (load + write only: needs no GPU).
 */
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "tmvs.h"

using namespace tmvs;

static bool loadAny(MVS &mvs, const std::string &file) {
    const size_t dot = file.find_last_of('.');
    const std::string ext = dot == std::string::npos ? "" : file.substr(dot + 1);
    if (ext == "nvm") return mvs.loadNVM(file.c_str(), false);
    if (ext == "nvm2") return mvs.loadNVM(file.c_str(), true);
    if (ext == "mvs") return mvs.loadMVS(file.c_str());
    fprintf(stderr, "unknown input type: %s\n", file.c_str());
    return false;
}

/* `tmvs -f file.mvs`, TMVS.cpp:124-170: PMVS cell / visibility / neighbour-cell filters, then the PCMVS neighbour
 * filter (its pair scan on the GPU), writing the reference's eight output files. */
static int runFiltering(MVS &mvs, const std::string &file, const std::string &outDir) {
    const size_t dot = file.find_last_of('.');
    if (dot == std::string::npos || file.substr(dot + 1) != "mvs") {
        printf("filtering only mvs file\n");
        return 1;
    }
    printf("patches: %zu\n", mvs.patches.size());
    typedef std::chrono::steady_clock Clock;
    const Clock::time_point t0 = Clock::now();
    double tFilter = 0, tWrite = 0;
    Clock::time_point t = t0;
    auto lap = [&](double &acc) { const Clock::time_point n = Clock::now(); acc += std::chrono::duration<double>(n - t).count(); t = n; };
    mvs.cellFiltering();
    lap(tFilter);
    mvs.writeMVS((outDir + "PMVS_filter1.mvs").c_str());
    mvs.writePLY((outDir + "PMVS_filter1.ply").c_str());
    lap(tWrite);
    mvs.visibilityFiltering();
    lap(tFilter);
    mvs.writeMVS((outDir + "PMVS_filter2.mvs").c_str());
    mvs.writePLY((outDir + "PMVS_filter2.ply").c_str());
    lap(tWrite);
    mvs.neighborCellFiltering(0.25);
    lap(tFilter);
    mvs.writeMVS((outDir + "PMVS_filter3.mvs").c_str());
    mvs.writePLY((outDir + "PMVS_filter3.ply").c_str());
    mvs.writeDeletedPatchMVS((outDir + "PMVS_filter_deleted.mvs").c_str());
    mvs.writeDeletedPatchPLY((outDir + "PMVS_filter_deleted.ply").c_str());
    mvs.clearDeletedPatches();
    lap(tWrite);
    printf("phase seconds: PMVS filters %.3f output %.3f\n", tFilter, tWrite);
    if (!mvs.neighborPatchFiltering(0.25)) { fprintf(stderr, "PCMVS filter failed: %s\n", mvs.lastError().c_str()); return 1; }
    double tPair = 0;
    lap(tPair);
    mvs.writeMVS((outDir + "PCMVS_filter.mvs").c_str());
    mvs.writePLY((outDir + "PCMVS_filter.ply").c_str());
    mvs.writeDeletedPatchMVS((outDir + "PCMVS_filter_deleted.mvs").c_str());
    mvs.writeDeletedPatchPLY((outDir + "PCMVS_filter_deleted.ply").c_str());
    lap(tWrite);
    printf("phase seconds: PCMVS filter %.3f (incl. CUDA context) output %.3f (all eight files)\n", tPair, tWrite);
    printf("patches kept: %zu\n", mvs.patches.size());
    printf("time1\t%f\n", std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
    return 0;
}

int main(int argc, char **argv) {
    std::string mode, input, configFile = "config.txt", outDir, imageDir, convertOut;
    int roundSize = 1024, device = 0, gpus = 1;
    unsigned long long seed = 42;
    double autosave = 5.0;
    bool expand = true, verbose = false, mergeSlots = true, pipeline = true;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        if ((a == "-r" || a == "-f" || a == "-v" || a == "-a") && i + 1 < argc) { mode = a; input = argv[++i]; }
        else if (a == "--convert" && i + 2 < argc) { mode = a; input = argv[++i]; convertOut = argv[++i]; }
        else if (a == "--config" && i + 1 < argc) configFile = argv[++i];
        else if (a == "--out-dir" && i + 1 < argc) outDir = argv[++i];
        else if (a == "--image-dir" && i + 1 < argc) imageDir = argv[++i];
        else if (a == "--round" && i + 1 < argc) roundSize = atoi(argv[++i]);
        else if (a == "--device" && i + 1 < argc) device = atoi(argv[++i]);
        else if (a == "--gpus" && i + 1 < argc) gpus = atoi(argv[++i]);
        else if (a == "--seed" && i + 1 < argc) seed = strtoull(argv[++i], nullptr, 10);
        else if (a == "--autosave-seconds" && i + 1 < argc) autosave = atof(argv[++i]);
        else if (a == "--no-expand") expand = false;
        else if (a == "--merge-slots") mergeSlots = true;
        else if (a == "--slot-passes") mergeSlots = false;
        else if (a == "--no-pipeline") pipeline = false;
        else if (a == "-V") verbose = true;
        else { fprintf(stderr, "unknown argument %s\n", a.c_str()); return 2; }
    }
    if (mode.empty()) {   /* TMVS.cpp:183-197 */
        printf("usage: tmvs -f <input.mvs> [--config config.txt] [--out-dir DIR] [--device D] [--gpus N]\n");
        printf("usage: tmvs -r <input.nvm|input.nvm2|input.mvs> [--config config.txt] [--out-dir DIR] [--round K] [--device D] [--gpus N] [--seed S] [--slot-passes] [--no-pipeline]\n");
        return 2;
    }
    if (mode == "-v" || mode == "-a") {
        fprintf(stderr, "%s is outside the scope of this build (reconstruction and filtering only)\n", mode.c_str());
        return 2;
    }
    if (!outDir.empty() && outDir[outDir.size() - 1] != '/') outDir += "/";

    if (!imageDir.empty() && imageDir[imageDir.size() - 1] != '/') imageDir += "/";

    MvsConfig config;
    setInitConfig(config);                                   /* TMVS.cpp:177 */
    loadConfig(configFile.c_str(), config);                  /* TMVS.cpp:178 */
    MVS mvs(config);                                         /* TMVS.cpp:181 */
    mvs.roundSize = roundSize > 0 ? roundSize : 1;
    mvs.device = device;
    mvs.numGpus = gpus > 0 ? gpus : 1;
    mvs.rngSeed = seed;
    mvs.autosaveSeconds = autosave;
    mvs.verbose = verbose;
    mvs.outDir = outDir;
    mvs.mergeSlots = mergeSlots;
    mvs.pipelineRounds = pipeline;
    mvs.imageDir = imageDir;

    if (!loadAny(mvs, input)) { fprintf(stderr, "load failed: %s\n", mvs.lastError().c_str()); return 1; }
    loadConfig(configFile.c_str(), config);                  /* TMVS.cpp:92-93: config.txt wins over the MVS header */
    mvs.setConfig(config);
    printf("cameras: %zu patches: %zu\n", mvs.cameras.size(), mvs.patches.size());
    if (mode == "--convert") {   /* writer by extension: MVS_V3 (default), .ply, .psr */
        const size_t dot = convertOut.find_last_of('.');
        const std::string ext = dot == std::string::npos ? "" : convertOut.substr(dot + 1);
        if (ext == "ply") return mvs.writePLY(convertOut.c_str()) ? 0 : 1;
        if (ext == "psr") return mvs.writePSR(convertOut.c_str()) ? 0 : 1;
        return mvs.writeMVS(convertOut.c_str()) ? 0 : 1;
    }
    if (mode == "-f") return runFiltering(mvs, input, outDir);
    if (mvs.patches.empty()) {
        fprintf(stderr, "no seed points in the input (SIFT seed generation, featuremanager.cpp, is out of scope)\n");
        return 1;
    }

    typedef std::chrono::steady_clock Clock;
    const Clock::time_point t0 = Clock::now();
    mvs.writeMVS((outDir + "init.mvs").c_str());
    if (!mvs.refineSeedPatches()) { fprintf(stderr, "seed refinement failed: %s\n", mvs.lastError().c_str()); return 1; }
    printf("seeds kept: %zu\n", mvs.patches.size());
    mvs.writeMVS((outDir + "seed.mvs").c_str());
    const Clock::time_point t1 = Clock::now();
    if (expand && !mvs.expansionPatches()) { fprintf(stderr, "expansion failed: %s\n", mvs.lastError().c_str()); return 1; }
    const Clock::time_point t2 = Clock::now();
    mvs.writeMVS((outDir + "exp.mvs").c_str());
    mvs.writePLY((outDir + "exp.ply").c_str());
    mvs.writePSR((outDir + "exp.psr").c_str());
    const Clock::time_point t3 = Clock::now();
    const double total = std::chrono::duration<double>(t3 - t0).count();
    printf("patches: %zu refined: %ld gpu_seconds: %f\n", mvs.patches.size(), mvs.refinedCount, mvs.gpuSeconds);
    if (mvs.refinedCount > 0)
        printf("per refined patch: evaluations %.1f (window loop %.1f) iterations %.1f runs %.2f views kept %.2f dropped %.1f %% LOD 0/1/2/3+ %ld/%ld/%ld/%ld\n",
               (double)mvs.statEvaluations / mvs.refinedCount, (double)mvs.statWindowEvaluations / mvs.refinedCount,
               (double)mvs.statIterations / mvs.refinedCount, (double)mvs.statRuns / mvs.refinedCount, (double)mvs.statViews / mvs.refinedCount,
               100.0 * mvs.statDropped / mvs.refinedCount, mvs.statLOD[0], mvs.statLOD[1], mvs.statLOD[2],
               mvs.statLOD[3] + mvs.statLOD[4] + mvs.statLOD[5] + mvs.statLOD[6] + mvs.statLOD[7]);
    printf("phase seconds: seeds %.3f (context %.3f) expansion %.3f output %.3f\n", std::chrono::duration<double>(t1 - t0).count(),
           mvs.contextSeconds, std::chrono::duration<double>(t2 - t1).count(), std::chrono::duration<double>(t3 - t2).count());
    printf("time1\t%f\n", total);                            /* TMVS.cpp:118-119 */
    return 0;
}
