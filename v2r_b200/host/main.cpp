// main.cpp -- `v2r idw ...` with the reference's flags (cmd/idw.go:50-67, cmd/root.go:34-36), GPU back end.
//
//   v2r idw -f points.txt [--epsg 2284] [--es 1.5 --ee 3.0 --ei 0.5] [-c --cx 200 --cy 200] [--outPath data/idw/]
//   v2r idw -g -f in.gpkg --layer wsels --field elevation --sx 100 --sy 100 ...
//   additive flags: --gpus N (1,2,4,8)  --fp32  --plan (print the grid / sweep plan and exit; needs no GPU)
//   v2r translate SRC.tiff DST.asc Int16        (processing.TransferType, used by the golden test)
//
// Flag semantics kept from the reference: --ee defaults to es+ei when not given (cmd/idw.go:80-82); ei <= 0 and
// ee <= es are fatal (:83-88); --sx/--sy only take effect with -g (txt mode reads STEP from the file) but always
// appear in the output name `{outPath}{stem}_step{sx}-{sy}[chunked]pow{exp:.1f}.tiff` (:147-155).
// The whole exponent sweep is ONE solve (one pass over the points) instead of one goroutine per exponent.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "v2r_host.hpp"

using namespace v2r;

namespace {
int g_level = 1;  // 0 = ERROR, 1 = INFO, 2 = DEBUG  (cmd/root.go:39-49)
void logf(int lvl, const char *tag, const std::string &m) {
    if (lvl <= g_level) fprintf(stderr, "[%s] %s\n", tag, m.c_str());
}
void info(const std::string &m) { logf(1, "INFO", m); }
void debug(const std::string &m) { logf(2, "DEBUG", m); }
[[noreturn]] void fatal(const std::string &m) {
    fprintf(stderr, "[FATAL] %s\n", m.c_str());
    exit(1);
}

struct Flags {
    bool fromGPKG = false, useChunking = false, fp32 = false, plan = false;
    long dumpPoints = 0;  // --dump-points N: print the first N parsed points (%.17g) with --plan; a test hook for the readers
    long idwChunkX = 200, idwChunkY = 200, epsg = 2284, gpus = 1;
    double expIncrement = .5, expStart = 1.5, expEnd = 0.0, stepX = 100.0, stepY = 100.0;
    bool eeChanged = false;
    std::string infile, layer, field, outfileFolder = "data/idw/";
};

std::string fmt(const char *f, double v) {
    char b[64];
    snprintf(b, sizeof b, f, v);
    return b;
}

int run_idw(int argc, char **argv) {
    Flags F;
    std::map<std::string, std::string> alias = {{"-g", "--gpkg"}, {"-c", "--concurrent"}, {"-f", "--file"}, {"-d", "--debug"}, {"-e", "--error"}, {"-l", "--log"}};
    for (int i = 0; i < argc; ++i) {
        std::string a = argv[i], val;
        bool has_val = false;
        size_t eq = a.find('=');
        if (a.rfind("--", 0) == 0 && eq != std::string::npos) { val = a.substr(eq + 1); a = a.substr(0, eq); has_val = true; }
        if (alias.count(a)) a = alias[a];
        auto need = [&]() -> std::string {
            if (has_val) return val;
            if (i + 1 >= argc) fatal("flag needs an argument: " + a);
            return argv[++i];
        };
        auto boolean = [&]() { return !has_val || val == "true" || val == "1"; };
        if (a == "--gpkg") F.fromGPKG = boolean();
        else if (a == "--concurrent") F.useChunking = boolean();
        else if (a == "--fp32") F.fp32 = boolean();
        else if (a == "--plan") F.plan = boolean();
        else if (a == "--dump-points") F.dumpPoints = atol(need().c_str());
        else if (a == "--debug") { if (boolean()) g_level = 2; }
        else if (a == "--error") { if (boolean() && g_level != 2) g_level = 0; }
        else if (a == "--log") { (void)boolean(); }  // lumberjack file logging: host plumbing, not reproduced
        else if (a == "--cx") F.idwChunkX = atol(need().c_str());
        else if (a == "--cy") F.idwChunkY = atol(need().c_str());
        else if (a == "--epsg") F.epsg = atol(need().c_str());
        else if (a == "--gpus") F.gpus = atol(need().c_str());
        else if (a == "--ei") F.expIncrement = atof(need().c_str());
        else if (a == "--es") F.expStart = atof(need().c_str());
        else if (a == "--ee") { F.expEnd = atof(need().c_str()); F.eeChanged = true; }
        else if (a == "--sx") F.stepX = atof(need().c_str());
        else if (a == "--sy") F.stepY = atof(need().c_str());
        else if (a == "--file") F.infile = need();
        else if (a == "--outPath") F.outfileFolder = need();
        else if (a == "--layer") F.layer = need();
        else if (a == "--field") F.field = need();
        else fatal("unknown flag: " + a);
    }
    // reqFlagsIDW / expRange (cmd/idw.go:71-90)
    if (F.fromGPKG && (F.layer.empty() || F.field.empty())) fatal("required flag(s) \"layer\", \"field\" not set");
    info("exp");
    if (!F.eeChanged) F.expEnd = F.expStart + F.expIncrement;
    if (F.expIncrement <= 0) fatal("exponential increment (" + fmt("%g", F.expIncrement) + ") must be > 0");
    if (F.expEnd <= F.expStart) fatal("exponent range: [" + fmt("%g", F.expStart) + ", " + fmt("%g", F.expEnd) + ") has no iterations");
    if (F.infile.empty()) fatal("required flag(s) \"file\" not set");

    info("IDW Started");
    info("-----Flags-----");
    info("Filepath: " + F.infile);
    info("Outfile folder: " + F.outfileFolder);
    info(std::string("Concurrent: ") + (F.useChunking ? "true" : "false"));
    info(std::string("From GPKG: ") + (F.fromGPKG ? "true" : "false"));
    if (F.fromGPKG) { info("Layer name: " + F.layer); info("Field name: " + F.field); }
    else info("Used epsg = " + std::to_string(F.epsg) + " on text file");
    info("Step size in x-direction: " + fmt("%g", F.stepX));
    info("Step size in y-direction: " + fmt("%g", F.stepY));
    info("Exponent: [" + fmt("%g", F.expStart) + ", " + fmt("%g", F.expEnd) + ")   Step size: " + fmt("%g", F.expIncrement));
    info("GPUs: " + std::to_string(F.gpus) + (F.fp32 ? "   precision: FP32 fast path" : "   precision: FP64"));
    info("---------------");

    const auto start = std::chrono::steady_clock::now();
    try {
        // doIDW (cmd/idw.go:117-170)
        processing::ReadResult rd = F.fromGPKG ? processing::ReadGeoPackage(F.infile, F.layer, F.field, F.stepX, F.stepY)
                                               : processing::ReadTextData(F.infile, (int)F.epsg);
        debug("projection " + rd.proj);
        tools::PointSet data = tools::MakeCoordSpace(rd.points, rd.xInfo, rd.yInfo);
        std::string stem = F.infile;
        size_t slash = stem.find_last_of("/\\");
        if (slash != std::string::npos) stem = stem.substr(slash + 1);
        size_t dot = stem.find_last_of('.');
        if (dot == std::string::npos) fatal("infile has no extension");  // the reference slices [:LastIndex(".")] and panics
        stem = stem.substr(0, dot);
        std::string outfile = F.outfileFolder + stem + "_step" + fmt("%.0f", F.stepX) + "-" + fmt("%.0f", F.stepY) + (F.useChunking ? "chunked" : "");
        debug(outfile);
        std::vector<double> pows = tools::ExpRange(F.expStart, F.expEnd, F.expIncrement);
        std::vector<std::string> outs;
        for (double p : pows) outs.push_back(outfile + "pow" + fmt("%.1f", p) + ".tiff");
        int64_t rows, cols;
        tools::GetDimensions(rd.xInfo, rd.yInfo, rows, cols);
        info("RowsXCols: [" + std::to_string(rows) + " X " + std::to_string(cols) + "]   points: " + std::to_string(rd.points.size()) +
             " -> " + std::to_string(data.size()) + " after MakeCoordSpace");
        if (F.plan) {
            char buf[512];
            double ops = 0;
            if (v2r_idw_describe(pows.data(), (int)pows.size(), F.fp32 ? V2R_IDW_FP32 : V2R_IDW_FP64, buf, sizeof buf, &ops) != V2R_IDW_OK)
                fatal(v2r_idw_last_error());
            printf("{\"rows\": %lld, \"cols\": %lld, \"points_in\": %lld, \"points\": %lld, \"exponents\": %zu, \"kernel\": \"%s\", \"first_outfile\": \"%s\"",
                   (long long)rows, (long long)cols, (long long)rd.points.size(), (long long)data.size(), pows.size(), buf, outs[0].c_str());
            if (F.dumpPoints > 0) {  // the raw reader output (before MakeCoordSpace), then the keyed set
                printf(", \"raw\": [");
                for (int64_t i = 0; i < rd.points.size() && i < F.dumpPoints; ++i)
                    printf("%s[%.17g, %.17g, %.17g]", i ? ", " : "", rd.points.x[i], rd.points.y[i], rd.points.w[i]);
                printf("], \"keyed\": [");
                for (int64_t i = 0; i < data.size() && i < F.dumpPoints; ++i)
                    printf("%s[%.17g, %.17g, %.17g, %lld, %lld]", i ? ", " : "", data.x[i], data.y[i], data.w[i], (long long)data.key_r[i], (long long)data.key_c[i]);
                printf("]");
            }
            printf("}\n");
            return 0;
        }
        idw::Options opt;
        opt.n_gpus = (int)F.gpus;
        opt.fp32 = F.fp32;
        opt.progress = [](const std::string &m) { info(m); };
        int64_t band = 0, chunkC = 0;
        if (F.useChunking) {  // ChunkSolve semantics: --cy x --cx is the write window (idw_chunk.go:48-53 for oversize chunks)
            band = F.idwChunkY;
            chunkC = F.idwChunkX;
            if (F.idwChunkY > rows || F.idwChunkX > cols) {
                logf(1, "WARN", "chunk size too large for total image; Chunk sizes changed to ~1/16 of total size");
                band = rows / 4;
                chunkC = cols / 4;
            }
            if (band <= 0 || chunkC <= 0) fatal("chunk size is not usable for this grid");
        }
        int iterations = 0;
        idw::SweepSolve(data, outs, rd.xInfo, rd.yInfo, rd.proj, pows, band, [&](const std::string &m) { ++iterations; debug(m); }, opt, chunkC);
        const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count();
        info("Completed " + std::to_string(iterations) + " iterations in " + fmt("%g", secs) + "s");
        info("Outfiles: " + outfile + "pow{EXP}.{EXT}");
        double h2d = 0, k = 0, d2h = 0;
        v2r_idw_last_timing(idw::Context(opt), &h2d, &k, &d2h);
        info("GPU: h2d " + fmt("%.2f", h2d) + " ms, kernels " + fmt("%.2f", k) + " ms, gather+d2h " + fmt("%.2f", d2h) + " ms, " +
             fmt("%.1f", k > 0 ? double(rows) * cols * data.size() / (k * 1e-3) / 1e9 : 0.0) + " G cell-point evals/s");
    } catch (const std::exception &e) {
        fatal(e.what());
    }
    info("IDW Finished");
    return 0;
}
}  // namespace

int main(int argc, char **argv) {
    if (argc >= 2 && !strcmp(argv[1], "idw")) return run_idw(argc - 2, argv + 2);
    if (argc == 5 && !strcmp(argv[1], "translate")) {
        try {
            processing::TransferType(argv[2], argv[3], argv[4]);
        } catch (const std::exception &e) {
            fatal(e.what());
        }
        return 0;
    }
    fprintf(stderr,
            "v2r is a library that implements Vector to Raster interpolation routines.\n"
            "This build carries the GPU (B200) IDW path only:\n"
            "  v2r idw -f FILE [-g --layer L --field F] [--sx --sy] [--es --ee --ei] [-c --cx --cy] [--epsg] [--outPath DIR/]\n"
            "          [--gpus N] [--fp32] [--plan] [-d|-e]\n"
            "  v2r translate SRC.tiff DST.asc Int16\n");
    return argc >= 2 && !strcmp(argv[1], "clean") ? (fprintf(stderr, "the cleaner is outside the scope of this build\n"), 2) : 1;
}
