// processing.cpp -- readers and raster sink of the C++ host (tools/processing of the reference).
//   ReadTextData    tools/processing/read_data.go:17-111  -- tokenizer appending straight into pinned SoA
//   ReadGeoPackage  tools/processing/read_data.go:149-173 + io_gdal.go:134-197 -- one sequential SQL scan decoding
//                   the GeoPackageBinary header itself (OGR's per-fid l.Feature(i) round trips are gone);
//                   libsqlite3 is loaded with dlopen (the image has libsqlite3.so.0 but no sqlite3.h)
//   WriteGDAL       tools/processing/io_gdal.go:27-74     -- native BigTIFF Float64 strip writer (GDAL is absent):
//                   create-or-update + window write at (offsets.C, offsets.R), south-up ModelTransformation
//   TransferType    tools/processing/io_gdal.go:115-132   -- `gdal_translate -ot Int16` to AAIGrid (idw_test.go:57)
#include <dlfcn.h>
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <regex>
#include <sstream>

#include "v2r_host.hpp"

namespace v2r {
namespace processing {

// ---- txt tokenizer ---------------------------------------------------------------------------------
// The whole file is read once and scanned in place (std::from_chars, ~25 ns per number); the reference's
// bufio.Scanner + strings.Fields + strconv.ParseFloat per line is what this replaces.
namespace {
struct Line { const char *b, *e; };
struct Tok { const char *b, *e; };

inline bool is_space(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\v' || c == '\f'; }

// next whitespace-separated token of [p, e)
inline bool next_tok(const char *&p, const char *e, Tok &t) {
    while (p < e && is_space(*p)) ++p;
    if (p >= e) return false;
    t.b = p;
    while (p < e && !is_space(*p)) ++p;
    t.e = p;
    return true;
}
// Grammar of strconv.ParseFloat (Go 1.18, base prefix only for hex): decimal floats, hex floats WITH a p exponent,
// [+-]inf / infinity and an unsigned nan in any case.  No underscores, no "nan(...)", no hex without exponent.
inline bool go_float_syntax(const char *b, const char *e) {
    auto ieq = [](const char *p, const char *q, const char *w) {
        for (; *w; ++p, ++w)
            if (p >= q || (*p | 0x20) != *w) return false;
        return p == q;
    };
    if (ieq(b, e, "nan")) return true;
    if (b < e && (*b == '+' || *b == '-')) ++b;
    if (ieq(b, e, "inf") || ieq(b, e, "infinity")) return true;
    bool hex = false;
    if (e - b >= 2 && b[0] == '0' && (b[1] | 0x20) == 'x') { hex = true; b += 2; }
    auto digit = [&](char c) { return (c >= '0' && c <= '9') || (hex && (c | 0x20) >= 'a' && (c | 0x20) <= 'f'); };
    int nd = 0;
    while (b < e && digit(*b)) { ++b; ++nd; }
    if (b < e && *b == '.') { ++b; while (b < e && digit(*b)) { ++b; ++nd; } }
    if (nd == 0) return false;
    if (b == e) return !hex;
    if ((*b | 0x20) != (hex ? 'p' : 'e')) return false;
    ++b;
    if (b < e && (*b == '+' || *b == '-')) ++b;
    if (b == e) return false;
    for (; b < e; ++b)
        if (*b < '0' || *b > '9') return false;
    return true;
}
// strconv.ParseFloat: ok == false on a syntax error (the reference then uses 0 for points, fails for STEP/ESTIMATE);
// an out-of-range value is +-Inf (Go returns it with ErrRange, which addPoints drops: the value is used)
inline double parse_float(Tok t, bool &ok) {
    ok = go_float_syntax(t.b, t.e);
    if (!ok) return 0.0;
    const char *b = t.b;
    if (*b == '+') ++b;
    double v = 0;
    auto r = std::from_chars(b, t.e, v);
    if (r.ec == std::errc() && r.ptr == t.e) return v;
    std::string s(t.b, t.e);  // hex floats and out-of-range values: strtod
    return strtod(s.c_str(), nullptr);
}
// strconv.Atoi(strings.TrimSpace(s)) with the error dropped (read_data.go:95): an optional sign and decimal digits,
// nothing else; anything that is not an integer gives 0 points
inline long long go_atoi_or_zero(const char *b, const char *e) {
    auto go_space = [](char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\v' || c == '\f'; };
    while (b < e && go_space(*b)) ++b;
    while (e > b && go_space(e[-1])) --e;
    const char *d = b;
    if (d < e && (*d == '+' || *d == '-')) ++d;
    if (d >= e) return 0;
    for (const char *q = d; q < e; ++q)
        if (*q < '0' || *q > '9') return 0;
    long long v = 0;
    auto r = std::from_chars(d, e, v);
    if (r.ec != std::errc()) return *b == '-' ? 0 : (long long)0x7fffffffffffffffLL;  // ErrRange: Atoi hands back the clamped value
    return *b == '-' ? -v : v;
}
// math.Min / math.Max of Go: NaN wins over every finite value, -Inf (+Inf) wins over NaN
inline double go_min(double a, double b) {
    if (std::isinf(a) && a < 0) return a;
    if (std::isinf(b) && b < 0) return b;
    if (std::isnan(a) || std::isnan(b)) return std::nan("");
    if (a == 0 && b == 0) return std::signbit(a) ? a : b;
    return a < b ? a : b;
}
inline double go_max(double a, double b) {
    if (std::isinf(a) && a > 0) return a;
    if (std::isinf(b) && b > 0) return b;
    if (std::isnan(a) || std::isnan(b)) return std::nan("");
    if (a == 0 && b == 0) return std::signbit(a) ? b : a;
    return a > b ? a : b;
}
inline double parse_strict(Tok t, const std::string &file) {
    bool ok;
    double v = parse_float(t, ok);
    if (!ok) throw Error(file + ": strconv.ParseFloat: parsing \"" + std::string(t.b, t.e) + "\": invalid syntax");
    return v;
}
}  // namespace

ReadResult ReadTextData(const std::string &filepath, int epsg) {
    std::string buf;
    {
        std::ifstream in(filepath, std::ios::binary);
        if (!in) throw Error("open " + filepath + ": no such file or directory");
        in.seekg(0, std::ios::end);
        const std::streamoff len = in.tellg();
        if (len < 0) throw Error("read " + filepath + ": is not a regular file");
        buf.resize((size_t)len);
        in.seekg(0);
        if (len > 0) in.read(&buf[0], (std::streamsize)buf.size());
    }
    ReadResult r;
    r.points = tools::PointSet(1024);
    r.proj = "EPSG:" + std::to_string(epsg);  // the reference asks GDAL for the WKT of the code (read_data.go:25-34)
    const char *p = buf.data(), *end = buf.data() + buf.size();
    auto next_line = [&](Line &l) {  // bufio.Scanner: lines split at \n, no token for a trailing empty line
        if (p >= end) return false;
        l.b = p;
        const char *nl = (const char *)memchr(p, '\n', (size_t)(end - p));
        l.e = nl ? nl : end;
        p = nl ? nl + 1 : end;
        return true;
    };
    Line l;
    while (next_line(l)) {
        const char *q = l.b;
        Tok head;
        if (!next_tok(q, l.e, head)) throw Error(filepath + ": blank line (the reference panics at read_data.go:40)");
        const size_t hl = (size_t)(head.e - head.b);
        if (hl == 6 && !memcmp(head.b, "POINTS", 6)) {
            // strconv.Atoi(strings.TrimSpace(strings.TrimPrefix(line, "POINTS "))), error dropped -> 0
            const char *c = (l.e - l.b >= 7 && !memcmp(l.b, "POINTS ", 7)) ? l.b + 7 : l.b;
            const long long n = go_atoi_or_zero(c, l.e);
            r.points = tools::PointSet(1024);  // `data = addPoints(sc)` (read_data.go:42): a later POINTS section replaces an earlier one
            for (long long i = 0; i < n; ++i) {
                Line pl;
                if (!next_line(pl)) pl.b = pl.e = end;
                const char *s = pl.b;
                Tok t[3];
                for (int k = 0; k < 3; ++k)
                    if (!next_tok(s, pl.e, t[k]))
                        throw Error(filepath + ": point line has fewer than 3 fields (the reference panics: index out of range)");
                bool ok;
                r.points.push_back(parse_float(t[0], ok), parse_float(t[1], ok), parse_float(t[2], ok));
            }
        } else if (hl == 4 && !memcmp(head.b, "STEP", 4)) {
            Line sl;
            if (!next_line(sl)) sl.b = sl.e = end;
            const char *s = sl.b;
            Tok a, b2;
            if (!next_tok(s, sl.e, a) || !next_tok(s, sl.e, b2)) throw Error(filepath + ": STEP needs two numbers");
            r.xInfo.Step = parse_strict(a, filepath);
            r.yInfo.Step = parse_strict(b2, filepath);
        } else if (hl == 8 && !memcmp(head.b, "ESTIMATE", 8)) {
            for (int xy = 0; xy < 2; ++xy) {
                Line el;
                if (!next_line(el)) el.b = el.e = end;
                const char *s = el.b;
                Tok t;
                for (int k = 0; next_tok(s, el.e, t); ++k) {
                    double val = parse_strict(t, filepath);
                    if (xy == 0 && k == 0) r.xInfo.Min = val;
                    if (xy == 0 && k == 1) r.xInfo.Max = val;
                    if (xy == 1 && k == 0) r.yInfo.Min = val;
                    if (xy == 1 && k == 1) r.yInfo.Max = val;
                }
            }
            return r;
        }
    }
    throw Error("ESTIMATE not in file");
}

// ---- sqlite3 through dlopen ------------------------------------------------------------------
namespace {
struct Sqlite {
    void *h = nullptr;
    int (*open_v2)(const char *, void **, int, const char *) = nullptr;
    int (*close)(void *) = nullptr;
    int (*prepare_v2)(void *, const char *, int, void **, const char **) = nullptr;
    int (*step)(void *) = nullptr;
    int (*finalize)(void *) = nullptr;
    int (*bind_text)(void *, int, const char *, int, void (*)(void *)) = nullptr;
    const void *(*column_blob)(void *, int) = nullptr;
    int (*column_bytes)(void *, int) = nullptr;
    double (*column_double)(void *, int) = nullptr;
    const unsigned char *(*column_text)(void *, int) = nullptr;
    long long (*column_int64)(void *, int) = nullptr;
    const char *(*errmsg)(void *) = nullptr;
};
Sqlite &sqlite() {
    static Sqlite s;
    if (s.h) return s;
    for (const char *nm : {"libsqlite3.so.0", "libsqlite3.so"}) {
        s.h = dlopen(nm, RTLD_NOW);
        if (s.h) break;
    }
    if (!s.h) throw Error(std::string("cannot load libsqlite3: ") + dlerror());
#define L(name) s.name = reinterpret_cast<decltype(s.name)>(dlsym(s.h, "sqlite3_" #name)); if (!s.name) throw Error("sqlite3_" #name " missing")
    L(open_v2); L(close); L(prepare_v2); L(step); L(finalize); L(bind_text); L(column_blob); L(column_bytes);
    L(column_double); L(column_text); L(column_int64); L(errmsg);
#undef L
    return s;
}
struct Db {
    void *db = nullptr;
    explicit Db(const std::string &path) {
        if (sqlite().open_v2(path.c_str(), &db, 1 /*READONLY*/, nullptr) != 0) {
            std::string m = db ? sqlite().errmsg(db) : "out of memory";
            if (db) sqlite().close(db);
            throw Error("open " + path + ": " + m);
        }
    }
    ~Db() { if (db) sqlite().close(db); }
};
struct Stmt {
    void *st = nullptr;
    Db &d;
    Stmt(Db &db, const std::string &sql) : d(db) {
        if (sqlite().prepare_v2(db.db, sql.c_str(), -1, &st, nullptr) != 0) throw Error(std::string("sql: ") + sqlite().errmsg(db.db) + " in " + sql);
    }
    ~Stmt() { if (st) sqlite().finalize(st); }
    void bind(int i, const std::string &v) { sqlite().bind_text(st, i, v.c_str(), -1, reinterpret_cast<void (*)(void *)>(-1)); }
    bool row() {
        int rc = sqlite().step(st);
        if (rc == 100) return true;
        if (rc == 101) return false;
        throw Error(std::string("sql step: ") + sqlite().errmsg(d.db));
    }
    std::string text(int c) { const unsigned char *t = sqlite().column_text(st, c); return t ? (const char *)t : ""; }
};
std::string quote_ident(const std::string &s) {
    std::string q = "\"";
    for (char c : s) { if (c == '"') q += '"'; q += c; }
    return q + "\"";
}
}  // namespace

ReadResult ReadGeoPackage(const std::string &filename, const std::string &layer, const std::string &field, double stepX, double stepY) {
    Db db(filename);
    {   // validGPKGLayer (io_gdal.go:134-159)
        Stmt s(db, "select table_name from gpkg_contents where data_type='features'");
        std::string all;
        bool ok = false;
        while (s.row()) { std::string t = s.text(0); ok |= (t == layer); all += (all.empty() ? "" : " ") + t; }
        if (!ok) throw Error("layer " + layer + " not in file. Available layers: [" + all + "]");
    }
    std::string geom_col;
    long long srs = 0;
    {
        Stmt s(db, "select column_name, srs_id from gpkg_geometry_columns where table_name=?1");
        s.bind(1, layer);
        if (!s.row()) throw Error("layer " + layer + " has no geometry column");
        geom_col = s.text(0);
        srs = sqlite().column_int64(s.st, 1);
    }
    ReadResult r;
    {
        Stmt s(db, "select definition from gpkg_spatial_ref_sys where srs_id=" + std::to_string(srs));
        if (s.row()) r.proj = s.text(0);
    }
    {
        Stmt s(db, "select name from pragma_table_info(?1)");
        s.bind(1, layer);
        bool ok = false;
        while (s.row()) ok |= (s.text(0) == field);
        if (!ok) throw Error("field " + field + " not in layer " + layer);
    }
    r.points = tools::PointSet(1024);
    r.xInfo.Step = stepX;
    r.yInfo.Step = stepY;
    Stmt s(db, "select " + quote_ident(geom_col) + ", " + quote_ident(field) + " from " + quote_ident(layer) + " order by rowid");
    // features in FID order like l.Feature(1..count) (io_gdal.go:180-190).  A GeoPackage feature table's primary key is
    // an INTEGER PRIMARY KEY, i.e. an alias of SQLite's rowid whatever its name is (fid, OBJECTID, id ...)
    while (s.row()) {
        const unsigned char *b = (const unsigned char *)sqlite().column_blob(s.st, 0);
        int nb = sqlite().column_bytes(s.st, 0);
        if (!b || nb < 8 || b[0] != 'G' || b[1] != 'P') throw Error("not a GeoPackageBinary geometry");
        static const int env_bytes[8] = {0, 32, 48, 48, 64, 0, 0, 0};
        int off = 8 + env_bytes[(b[3] >> 1) & 7];
        if (nb < off + 21) throw Error("truncated geometry blob");
        bool le = b[off] == 1;
        auto rd_u32 = [&](int o) { uint32_t v; memcpy(&v, b + o, 4); return le ? v : __builtin_bswap32(v); };
        auto rd_f64 = [&](int o) { uint64_t v; memcpy(&v, b + o, 8); if (!le) v = __builtin_bswap64(v); double d; memcpy(&d, &v, 8); return d; };
        if (rd_u32(off + 1) % 1000 != 1) throw Error("POINT geometries expected");
        double x = rd_f64(off + 5), y = rd_f64(off + 13);
        r.points.push_back(x, y, sqlite().column_double(s.st, 1));
        r.xInfo.Min = go_min(r.xInfo.Min, x); r.xInfo.Max = go_max(r.xInfo.Max, x);  // math.Min / math.Max (read_data.go:158-162)
        r.yInfo.Min = go_min(r.yInfo.Min, y); r.yInfo.Max = go_max(r.yInfo.Max, y);
    }
    return r;
}

// ---- raster sink ---------------------------------------------------------------------------------
static const long long kDataOffset = 4096;  // header + IFD + tag payloads live below it

static bool ends_with(const std::string &s, const char *suf) {
    size_t n = strlen(suf);
    return s.size() >= n && s.compare(s.size() - n, n, suf) == 0;
}

namespace {
struct Tag { uint16_t tag, type; uint64_t count; std::string payload; };
template <typename T> std::string bytes_of(const T &v) { return std::string((const char *)&v, sizeof v); }

int epsg_of(const std::string &wkt) {
    static const std::regex tail("AUTHORITY\\[\"EPSG\",\"(\\d+)\"\\]\\]\\s*$");
    std::smatch m;
    if (std::regex_search(wkt, m, tail)) return atoi(m[1].str().c_str());
    if (wkt.rfind("EPSG:", 0) == 0) return atoi(wkt.c_str() + 5);
    return -1;
}

void tiff_create(const std::string &path, const GDalInfo &info, long long rows, long long cols) {
    std::vector<Tag> tags;
    auto add_short = [&](uint16_t t, uint16_t v) { tags.push_back({t, 3, 1, bytes_of(v)}); };
    auto add_long8 = [&](uint16_t t, uint64_t v) { tags.push_back({t, 16, 1, bytes_of(v)}); };
    const uint64_t nbytes = (uint64_t)rows * (uint64_t)cols * 8;
    add_long8(256, (uint64_t)cols); add_long8(257, (uint64_t)rows); add_short(258, 64); add_short(259, 1); add_short(262, 1);
    add_long8(273, (uint64_t)kDataOffset); add_short(277, 1); add_long8(278, (uint64_t)std::max<long long>(rows, 1));
    add_long8(279, nbytes); add_short(284, 1); add_short(339, 3);
    double xf[16] = {info.XCell, 0, 0, info.XMin, 0, info.YCell, 0, info.YMin, 0, 0, 0, 0, 0, 0, 0, 1};
    tags.push_back({34264, 12, 16, std::string((const char *)xf, sizeof xf)});  // ModelTransformationTag (south-up)
    std::vector<uint16_t> gk = {1, 1, 0, 0, 1024, 0, 1, 1, 1025, 0, 1, 1};
    std::string ascii;
    int epsg = epsg_of(info.Proj);
    if (epsg > 0 && epsg < 65536) { gk.insert(gk.end(), {3072, 0, 1, (uint16_t)epsg}); }
    else if (!info.Proj.empty() && info.Proj.size() < 60000) {
        ascii = info.Proj + "|";
        ascii.push_back('\0');
        gk.insert(gk.end(), {3072, 0, 1, 32767, 3073, 34737, (uint16_t)(ascii.size() - 1), 0});
    }
    gk[3] = (uint16_t)((gk.size() - 4) / 4);
    tags.push_back({34735, 3, gk.size(), std::string((const char *)gk.data(), gk.size() * 2)});
    if (!ascii.empty()) tags.push_back({34737, 2, ascii.size(), ascii});
    std::sort(tags.begin(), tags.end(), [](const Tag &a, const Tag &b) { return a.tag < b.tag; });

    const uint64_t ifd_off = 16, ifd_size = 8 + 20 * tags.size() + 8;
    uint64_t extra_off = ifd_off + ifd_size;
    std::string entries, extra;
    for (const Tag &t : tags) {
        entries += bytes_of(t.tag) + bytes_of(t.type) + bytes_of(t.count);
        if (t.payload.size() <= 8) {
            std::string v = t.payload;
            v.resize(8, '\0');
            entries += v;
        } else {
            entries += bytes_of((uint64_t)(extra_off + extra.size()));
            extra += t.payload;
            extra.resize((extra.size() + 7) / 8 * 8, '\0');
        }
    }
    std::string head = "II";
    head += bytes_of((uint16_t)43) + bytes_of((uint16_t)8) + bytes_of((uint16_t)0) + bytes_of(ifd_off);
    head += bytes_of((uint64_t)tags.size()) + entries + bytes_of((uint64_t)0) + extra;
    if ((long long)head.size() > kDataOffset) throw Error("projection text too long for the reserved TIFF header block");
    head.resize((size_t)kDataOffset, '\0');
    int fd = ::open(path.c_str(), O_CREAT | O_TRUNC | O_WRONLY, 0644);
    if (fd < 0) throw Error("cannot create " + path + ": " + strerror(errno));
    bool ok = ::write(fd, head.data(), head.size()) == (ssize_t)head.size() && ::ftruncate(fd, kDataOffset + (off_t)nbytes) == 0;
    ::close(fd);
    if (!ok) throw Error("cannot write " + path);
}

struct TiffLayout { long long rows = 0, cols = 0, data_off = 0; double xf[16] = {0}; bool has_xf = false; };
TiffLayout tiff_layout(const std::string &path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw Error("cannot open " + path);
    std::string head((size_t)kDataOffset, '\0');
    f.read(&head[0], kDataOffset);
    if (f.gcount() < 16 || head[0] != 'I' || head[1] != 'I' || (uint8_t)head[2] != 43) throw Error(path + ": not a little-endian BigTIFF");
    auto u64 = [&](size_t o) { uint64_t v; memcpy(&v, &head[o], 8); return v; };
    auto u16 = [&](size_t o) { uint16_t v; memcpy(&v, &head[o], 2); return v; };
    const uint64_t have = (uint64_t)f.gcount();
    const uint64_t ifd = u64(8);
    if (ifd > have || have - ifd < 8) throw Error(path + ": IFD outside the header block");
    const uint64_t n = u64(ifd);
    if (n > (have - ifd - 8) / 20) throw Error(path + ": IFD outside the header block");
    TiffLayout L;
    for (uint64_t i = 0; i < n; ++i) {
        size_t e = ifd + 8 + 20 * i;
        uint16_t tag = u16(e);
        uint64_t val = u64(e + 12);
        if (tag == 256) L.cols = (long long)val;
        if (tag == 257) L.rows = (long long)val;
        if (tag == 273) L.data_off = (long long)val;
        if (tag == 34264) {
            if (val > have || have - val < 128 || u64(e + 4) != 16) throw Error(path + ": ModelTransformationTag outside the header block");
            memcpy(L.xf, &head[val], 128);
            L.has_xf = true;
        }
    }
    if (L.rows < 0 || L.cols < 0 || L.data_off < 0) throw Error(path + ": bad raster dimensions");
    return L;
}

std::string asc_header(long long cols, long long rows, double x_min, double y_ll, double cell) {
    char b[256];
    snprintf(b, sizeof b, "ncols        %lld\nnrows        %lld\nxllcorner    %.12f\nyllcorner    %.12f\ncellsize     %.12f\n", cols, rows, x_min, y_ll, cell);
    return b;
}
}  // namespace

void WriteGDAL(const double *buf, const GDalInfo &info, const std::string &filename, tools::OrderedPair offsets,
               tools::OrderedPair totalSize, tools::OrderedPair bufferSize, bool create) {
    const bool tiff = ends_with(filename, ".tif") || ends_with(filename, ".tiff"), asc = ends_with(filename, ".asc");
    if (!tiff && !asc) throw Error("Not a valid output file extension. Only accepts tif and ascii files (.tif/.tiff and .asc).");
    const long long rows = totalSize.R, cols = totalSize.C, r0 = offsets.R, c0 = offsets.C, br = bufferSize.R, bc = bufferSize.C;
    if (asc) {
        if (r0 != 0 || c0 != 0 || br != rows || bc != cols) throw Error(".asc output supports only whole-raster writes");
        FILE *f = fopen(filename.c_str(), "w");
        if (!f) throw Error("cannot create " + filename);
        fputs(asc_header(cols, rows, info.XMin, info.YMin - rows * info.YCell, info.XCell).c_str(), f);
        for (long long r = 0; r < rows; ++r) {
            for (long long c = 0; c < cols; ++c) fprintf(f, " %.17g", buf[r * cols + c]);
            fputc('\n', f);
        }
        fclose(f);
        return;
    }
    if (create) tiff_create(filename, info, rows, cols);
    TiffLayout L = tiff_layout(filename);
    if (r0 < 0 || c0 < 0 || r0 + br > L.rows || c0 + bc > L.cols) throw Error("Access window out of range in RasterIO()");
    int fd = ::open(filename.c_str(), O_WRONLY);
    if (fd < 0) throw Error("cannot open " + filename + " for update");
    auto put = [&](const double *p, size_t bytes, long long off) {
        const char *q = (const char *)p;
        while (bytes) {
            ssize_t w = ::pwrite(fd, q, bytes, (off_t)off);
            if (w <= 0) { ::close(fd); throw Error("write failed on " + filename); }
            q += w; off += w; bytes -= (size_t)w;
        }
    };
    if (c0 == 0 && bc == L.cols) put(buf, (size_t)br * bc * 8, L.data_off + r0 * L.cols * 8);
    else
        for (long long i = 0; i < br; ++i) put(buf + i * bc, (size_t)bc * 8, L.data_off + ((r0 + i) * L.cols + c0) * 8);
    ::close(fd);
}

void TransferType(const std::string &src, const std::string &dst, const std::string &outputType) {
    if (outputType != "Int16" || !ends_with(dst, ".asc")) throw Error("only -ot Int16 to .asc is supported");
    TiffLayout L = tiff_layout(src);
    if (!L.has_xf) throw Error(src + ": no geotransform");
    std::ifstream f(src, std::ios::binary);
    f.seekg(L.data_off);
    std::vector<double> row((size_t)L.cols);
    FILE *o = fopen(dst.c_str(), "w");
    if (!o) throw Error("cannot create " + dst);
    fputs(asc_header(L.cols, L.rows, L.xf[3], L.xf[7] - L.rows * L.xf[5], L.xf[0]).c_str(), o);
    for (long long r = 0; r < L.rows; ++r) {
        f.read((char *)row.data(), L.cols * 8);
        for (long long c = 0; c < L.cols; ++c) {
            double v = row[(size_t)c];
            double q = std::isnan(v) ? 0.0 : std::copysign(std::floor(std::fabs(v) + 0.5), v);  // round half away from zero
            q = std::fmin(std::fmax(q, -32768.0), 32767.0);
            fprintf(o, " %d", (int)q);
        }
        fputc('\n', o);
    }
    fclose(o);
}

}  // namespace processing
}  // namespace v2r
