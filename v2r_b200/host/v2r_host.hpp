// v2r_host.hpp -- C++ host side above the C ABI (include/v2r_idw.h), mirroring the reference's Go
// interface for the IDW path: same names, argument meaning and error behaviour.
//
//   tools::Point / OrderedPair / Info / MakeInfo      tools/utils.go:46-55,123-131
//   tools::GetDimensions / PointToPair                tools/utils.go:133-137, 70-72
//   tools::MakeCoordSpace                             tools/utils.go:97-119   (input-order, deterministic)
//   processing::ReadTextData / ReadGeoPackage         tools/processing/read_data.go:17-111,149-173
//   processing::GDalInfo / CreateGDalInfo / WriteGDAL tools/processing/io_gdal.go:13-74
//   processing::TransferType                          tools/processing/io_gdal.go:115-132 (Int16 AAIGrid only)
//   idw::FullSolve / ChunkSolve                       features/idw/idw.go:15, features/idw/idw_chunk.go:43
//   idw::SweepSolve                                   the exponent loop of cmd/idw.go:153-161 as ONE call
//
// The reference is Go; no Go toolchain exists in this image, so the compiled host is C++ (the cgo
// binding a v2r maintainer adds is go/features/idw/idw_gpu.go).  Points live in page-locked SoA
// buffers from v2r_idw_alloc_pinned from the moment the readers parse them.
#pragma once
#include <cstdint>
#include <functional>
#include <limits>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/v2r_idw.h"

namespace v2r {

struct Error : std::runtime_error {
    using std::runtime_error::runtime_error;
};

namespace tools {

struct Point { double X = 0, Y = 0, Weight = 0; };
struct OrderedPair { int64_t R = 0, C = 0; };
struct Info {
    double Min = std::numeric_limits<double>::infinity();
    double Max = -std::numeric_limits<double>::infinity();
    double Step = 1.0;
};
inline Info MakeInfo() { return Info{}; }

// Page-locked SoA: x, y, weight (+ keys after MakeCoordSpace).  Movable, not copyable.
class PointSet {
public:
    PointSet() = default;
    explicit PointSet(int64_t capacity, bool with_keys = false);
    PointSet(PointSet &&o) noexcept;
    PointSet &operator=(PointSet &&o) noexcept;
    PointSet(const PointSet &) = delete;
    PointSet &operator=(const PointSet &) = delete;
    ~PointSet();
    void push_back(double x, double y, double w);  // grows geometrically (readers append here)
    int64_t size() const { return n_; }
    double *x = nullptr, *y = nullptr, *w = nullptr;
    int64_t *key_r = nullptr, *key_c = nullptr;  // null until keyed
private:
    void reserve(int64_t cap);
    void release();
    int64_t n_ = 0, cap_ = 0;
    bool keys_ = false, pinned_ = false;
    friend PointSet MakeCoordSpace(const PointSet &, const Info &, const Info &);
};

void GetDimensions(const Info &xInfo, const Info &yInfo, int64_t &rows, int64_t &cols);
OrderedPair PointToPair(const Point &p, const Info &xInfo, const Info &yInfo);
PointSet MakeCoordSpace(const PointSet &listPoints, const Info &xInfo, const Info &yInfo);
std::vector<double> ExpRange(double es, double ee, double ei);  // throws like expRange (cmd/idw.go:83-88)

}  // namespace tools

namespace processing {

struct GDalInfo { double XMin, YMin, XCell, YCell; int GDalDataType; std::string Proj; };
inline GDalInfo CreateGDalInfo(double XMin, double YMin, double XCell, double YCell, int type, const std::string &proj) {
    return GDalInfo{XMin, YMin, XCell, YCell, type, proj};
}

struct ReadResult { tools::PointSet points; std::string proj; tools::Info xInfo, yInfo; };
ReadResult ReadTextData(const std::string &filepath, int epsg);
ReadResult ReadGeoPackage(const std::string &filename, const std::string &layer, const std::string &field, double stepX, double stepY);

// offsets / totalSize / bufferSize are (R, C) pairs like tools.OrderedPair
void WriteGDAL(const double *unwrappedMatrix, const GDalInfo &info, const std::string &filename, tools::OrderedPair offsets,
               tools::OrderedPair totalSize, tools::OrderedPair bufferSize, bool create);
void TransferType(const std::string &src, const std::string &dst, const std::string &outputType);

}  // namespace processing

namespace idw {

using Channel = std::function<void(const std::string &)>;  // `channel <- msg`
struct Options {
    int n_gpus = 1;
    bool fp32 = false;
    Channel progress;  // optional: receives ChunkSolve-style "~40%, 8 / 21" lines while bands stream out
};

// One context per process (lazy).  Throws v2r::Error with the library's message on failure.
v2r_idw_ctx *Context(const Options &opt);

void SweepSolve(const tools::PointSet &data, const std::vector<std::string> &outfiles, const tools::Info &xInfo,
                const tools::Info &yInfo, const std::string &proj, const std::vector<double> &pows, int64_t bandRows,
                const Channel &channel, const Options &opt = {}, int64_t chunkC = 0);
void FullSolve(const tools::PointSet &data, const std::string &outfile, const tools::Info &xInfo, const tools::Info &yInfo,
               const std::string &proj, double pow, const Channel &channel, const Options &opt = {});
void ChunkSolve(const tools::PointSet &data, const std::string &outfile, const tools::Info &xInfo, const tools::Info &yInfo,
                int64_t chunkR, int64_t chunkC, const std::string &proj, double pow, const Channel &channel,
                const Options &opt = {});

}  // namespace idw
}  // namespace v2r
