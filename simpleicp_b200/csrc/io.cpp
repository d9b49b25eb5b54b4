// io.cpp — .xyz point-cloud text I/O (host only).  SURVEY.md §8f rank 1: on the lidar sets file
// parsing (np.genfromtxt of 1.3 M lines: seconds) dwarfs the GPU time, so the reader is part of
// the library.  Format contract of the reference readers (c++/src/simpleicp-cli.cpp:72-128,
// rust/src/io.rs:9-37, np.genfromtxt in python/simpleicp/tests/test_simpleicp.py:102-103):
// whitespace-separated x y z per line; lines starting with '/' or '#' (the CloudCompare header
// the reference's write_xyz emits, python/simpleicp/pointcloud.py:219-226) and blank lines are
// skipped; extra columns are ignored.  Numbers are parsed with std::from_chars (correctly
// rounded, so the doubles equal what Python/NumPy parse from the same text).
#include <charconv>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <string>
#include <thread>
#include <vector>

#include "../../include/sicp_b200.h"

namespace {

thread_local std::string g_io_error;

struct Chunk {
  const char* b;
  const char* e;
  std::vector<double> xyz;
  long long bad_line = -1;
};

inline const char* skip_ws(const char* p, const char* e) {
  while (p < e && (*p == ' ' || *p == '\t' || *p == '\r' || *p == ',' || *p == ';')) ++p;
  return p;
}

void parse_chunk_impl(Chunk& c);

// runs inside std::thread: an exception escaping it would call std::terminate
void parse_chunk(Chunk& c) {
  try {
    parse_chunk_impl(c);
  } catch (...) {
    c.xyz.clear();
    c.bad_line = -2;  // out of memory
  }
}

void parse_chunk_impl(Chunk& c) {
  const char* p = c.b;
  long long line = 0;
  c.xyz.reserve((size_t)(c.e - c.b) / 24 * 3 + 3);
  while (p < c.e) {
    const char* eol = (const char*)memchr(p, '\n', (size_t)(c.e - p));
    if (!eol) eol = c.e;
    const char* q = skip_ws(p, eol);
    if (q < eol && *q != '/' && *q != '#') {
      double v[3];
      int k = 0;
      for (; k < 3; ++k) {
        q = skip_ws(q, eol);
        if (q < eol && *q == '+') ++q;
        auto r = std::from_chars(q, eol, v[k]);
        if (r.ec != std::errc()) break;
        q = r.ptr;
      }
      if (k == 3) {
        c.xyz.push_back(v[0]);
        c.xyz.push_back(v[1]);
        c.xyz.push_back(v[2]);
      } else if (c.bad_line < 0) {
        c.bad_line = line;
      }
    }
    p = eol + 1;
    ++line;
  }
}

}  // namespace

extern "C" {

const char* sicp_io_last_error(void) { return g_io_error.c_str(); }

static int32_t xyz_load_impl(const char* path, double** xyz, int64_t* n);

int32_t sicp_xyz_load(const char* path, double** xyz, int64_t* n) {
  if (!path || !xyz || !n) {
    g_io_error = "NULL argument";
    return SICP_ERR_BAD_ARG;
  }
  *xyz = nullptr;
  *n = 0;
  try {
    return xyz_load_impl(path, xyz, n);
  } catch (const std::exception& e) {  // nothing may unwind through the C boundary
    g_io_error = std::string("reading ") + path + ": " + e.what();
    return SICP_ERR_BAD_ARG;
  }
}

static int32_t xyz_load_impl(const char* path, double** xyz, int64_t* n) {
  FILE* f = fopen(path, "rb");
  if (!f) {
    g_io_error = std::string("cannot open ") + path;
    return SICP_ERR_BAD_ARG;
  }
  fseek(f, 0, SEEK_END);
  const long long size = ftell(f);
  fseek(f, 0, SEEK_SET);
  std::vector<char> buf((size_t)size + 1);
  const size_t got = fread(buf.data(), 1, (size_t)size, f);
  fclose(f);
  if ((long long)got != size) {
    g_io_error = std::string("short read on ") + path;
    return SICP_ERR_BAD_ARG;
  }
  buf[(size_t)size] = '\n';
  unsigned hw = std::thread::hardware_concurrency();
  int nt = (int)std::min<long long>(std::max(1u, std::min(hw, 32u)), std::max<long long>(1, size / (1 << 20)));
  std::vector<Chunk> chunks((size_t)nt);
  const char* base = buf.data();
  const char* end = base + size;
  const char* cur = base;
  for (int t = 0; t < nt; ++t) {
    const char* stop = (t == nt - 1) ? end : base + size * (t + 1) / nt;
    if (stop < cur) stop = cur;  // a line longer than a whole share: the earlier chunk took it
    while (stop < end && *stop != '\n') ++stop;
    if (stop < end) ++stop;
    chunks[(size_t)t].b = cur;
    chunks[(size_t)t].e = stop;
    cur = stop;
  }
  std::vector<std::thread> th;
  for (int t = 1; t < nt; ++t) th.emplace_back(parse_chunk, std::ref(chunks[(size_t)t]));
  parse_chunk(chunks[0]);
  for (auto& x : th) x.join();
  size_t total = 0;
  for (auto& c : chunks) {
    if (c.bad_line == -2) {
      g_io_error = std::string("out of memory while parsing ") + path;
      return SICP_ERR_BAD_ARG;
    }
    if (c.bad_line >= 0) {
      g_io_error = std::string("malformed line in ") + path + " (need three numbers per line)";
      return SICP_ERR_BAD_ARG;
    }
    total += c.xyz.size();
  }
  double* out = (double*)malloc(std::max<size_t>(total, 1) * sizeof(double));
  if (!out) {
    g_io_error = "out of memory";
    return SICP_ERR_BAD_ARG;
  }
  size_t off = 0;
  for (auto& c : chunks) {
    if (!c.xyz.empty()) memcpy(out + off, c.xyz.data(), c.xyz.size() * sizeof(double));
    off += c.xyz.size();
  }
  *xyz = out;
  *n = (int64_t)(total / 3);
  return SICP_OK;
}

void sicp_xyz_free(double* xyz) { free(xyz); }

// CloudCompare-style text file like the reference's PointCloud.write_xyz ("//X Y Z" header,
// %.3f by default there).  decimals < 0 writes round-trip precision (%.17g).
int32_t sicp_xyz_save(const char* path, const double* xyz, int64_t n, int32_t decimals,
                      int32_t header) {
  if (!path || (!xyz && n > 0)) {
    g_io_error = "NULL argument";
    return SICP_ERR_BAD_ARG;
  }
  FILE* f = fopen(path, "wb");
  if (!f) {
    g_io_error = std::string("cannot open ") + path + " for writing";
    return SICP_ERR_BAD_ARG;
  }
  std::vector<char> buf(1 << 22);
  setvbuf(f, buf.data(), _IOFBF, buf.size());
  if (header) fputs("//X Y Z\n", f);
  char fmt[64];
  if (decimals < 0)
    snprintf(fmt, sizeof(fmt), "%%.17g %%.17g %%.17g\n");
  else
    snprintf(fmt, sizeof(fmt), "%%.%df %%.%df %%.%df\n", decimals, decimals, decimals);
  for (int64_t i = 0; i < n; ++i) fprintf(f, fmt, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
  const bool ok = (fclose(f) == 0);
  if (!ok) g_io_error = std::string("write error on ") + path;
  return ok ? SICP_OK : SICP_ERR_BAD_ARG;
}

}  // extern "C"
