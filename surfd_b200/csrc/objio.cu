// objio.cu -- host-side Wavefront .obj writer / reader of the output stage (no device code).
//
// The reference writes every mesh through open3d (`o3d.io.write_triangle_mesh`, utils/utils.py:79-121 +
// sample/generate_uncond.py:113-116) and re-reads / re-writes it through pymeshlab (generate_uncond.py:117-122): C++ on both
// sides.  A 512^3 shape has 0.34 M vertices and 0.68 M faces (1.3 M faces on the --watertight branch); formatting those in
// Python takes longer than generating the shape (2-4.5 s per file against 0.3 s of GPU time), so the text conversion lives
// here.  The layouts (headers, separators, number formats) are decided by the callers in surfd_b200/output.py.
#include <cerrno>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.cuh"

namespace surfd {

// printf("%f") / ("%g") as Python's % operator prints them: identical to C except that a NaN never carries a sign
static inline int fmt_double(char* out, double x, int general) {
  if (std::isnan(x)) { memcpy(out, "nan", 3); return 3; }
  return general ? snprintf(out, 40, "%g", x) : snprintf(out, 400, "%f", x);
}

static inline int fmt_int(char* out, long long v) {
  char tmp[24];
  int n = 0;
  bool neg = v < 0;
  unsigned long long u = neg ? (unsigned long long)(-(v + 1)) + 1ull : (unsigned long long)v;
  do { tmp[n++] = (char)('0' + u % 10); u /= 10; } while (u);
  int k = 0;
  if (neg) out[k++] = '-';
  while (n) out[k++] = tmp[--n];
  return k;
}

}  // namespace surfd

// header + nv lines "v x y z" + mid + nf lines "f a b c" (indices written 1-based) + footer.
// number_format: 0 = "%f" (MeshLab exporter), 1 = "%g" (C++ ostream default, open3d).  verts: host double [nv][3],
// faces: host int64 [nf][3], 0-based.
extern "C" int surfd_obj_write(const char* path, const char* header, const double* verts, int64_t nv, int number_format,
                               const char* mid, const int64_t* faces, int64_t nf, const char* footer) {
  using namespace surfd;
  SURFD_REQUIRE(path && (nv == 0 || verts) && (nf == 0 || faces) && nv >= 0 && nf >= 0, "bad argument");
  SURFD_REQUIRE(number_format == 0 || number_format == 1, "number_format must be 0 (%f) or 1 (%g)");
  FILE* fh = fopen(path, "wb");
  if (!fh) return set_error(SURFD_BAD_ARGUMENT, strerror(errno), __FILE__, __LINE__);
  std::vector<char> buf(1 << 20);
  const size_t cap = buf.size();
  size_t used = 0;
  bool ok = true;
  auto flush = [&]() { if (used) { ok = ok && fwrite(buf.data(), 1, used, fh) == used; used = 0; } };
  if (header) ok = ok && fputs(header, fh) >= 0;
  for (int64_t i = 0; i < nv; ++i) {
    if (used + 1400 > cap) flush();      // a "%f" of a huge double can take ~320 characters
    char* p = buf.data() + used;
    *p++ = 'v';
    for (int c = 0; c < 3; ++c) { *p++ = ' '; p += fmt_double(p, verts[3 * i + c], number_format); }
    *p++ = '\n';
    used = (size_t)(p - buf.data());
  }
  flush();
  if (mid) ok = ok && fputs(mid, fh) >= 0;
  for (int64_t i = 0; i < nf; ++i) {
    if (used + 128 > cap) flush();
    char* p = buf.data() + used;
    *p++ = 'f';
    for (int c = 0; c < 3; ++c) { *p++ = ' '; p += fmt_int(p, (long long)faces[3 * i + c] + 1); }
    *p++ = '\n';
    used = (size_t)(p - buf.data());
  }
  flush();
  if (footer) ok = ok && fputs(footer, fh) >= 0;
  ok = (fclose(fh) == 0) && ok;
  if (!ok) return set_error(SURFD_BAD_ARGUMENT, "short write", __FILE__, __LINE__);
  return 0;
}

// Minimal reader (`v x y z` and `f a[/..] b[/..] c[/..]` records, 1-based indices; everything else is skipped).
// Call with verts == faces == NULL to get the counts, then with buffers of those sizes (host double [nv][3], int64 [nf][3],
// 0-based on return).
extern "C" int surfd_obj_read(const char* path, int64_t* nv, int64_t* nf, double* verts, int64_t* faces) {
  using namespace surfd;
  SURFD_REQUIRE(path && nv && nf, "null argument");
  FILE* fh = fopen(path, "rb");
  if (!fh) return set_error(SURFD_BAD_ARGUMENT, strerror(errno), __FILE__, __LINE__);
  fseek(fh, 0, SEEK_END);
  const long size = ftell(fh);
  fseek(fh, 0, SEEK_SET);
  std::vector<char> data((size_t)size + 1);
  const bool ok = size == 0 || fread(data.data(), 1, (size_t)size, fh) == (size_t)size;
  fclose(fh);
  if (!ok) return set_error(SURFD_BAD_ARGUMENT, "short read", __FILE__, __LINE__);
  data[(size_t)size] = '\0';
  const bool fill = verts != nullptr || faces != nullptr;
  const int64_t cap_v = *nv, cap_f = *nf;
  int64_t cv = 0, cf = 0;
  char* p = data.data();
  char* end = p + size;
  while (p < end) {
    char* eol = (char*)memchr(p, '\n', (size_t)(end - p));
    if (!eol) eol = end;
    if (eol - p >= 2 && p[1] == ' ' && (p[0] == 'v' || p[0] == 'f')) {
      const char saved = *eol;
      *eol = '\0';
      if (p[0] == 'v') {
        if (fill) {
          if (cv >= cap_v || !verts) return set_error(SURFD_BAD_ARGUMENT, "vertex buffer too small", __FILE__, __LINE__);
          char* q = p + 2;
          for (int c = 0; c < 3; ++c) {
            char* next = nullptr;
            verts[3 * cv + c] = strtod(q, &next);
            if (next == q) return set_error(SURFD_BAD_ARGUMENT, "malformed vertex record", __FILE__, __LINE__);
            q = next;
          }
        }
        ++cv;
      } else {
        if (fill) {
          if (cf >= cap_f || !faces) return set_error(SURFD_BAD_ARGUMENT, "face buffer too small", __FILE__, __LINE__);
          char* q = p + 2;
          for (int c = 0; c < 3; ++c) {
            char* next = nullptr;
            const long long v = strtoll(q, &next, 10);
            if (next == q) return set_error(SURFD_BAD_ARGUMENT, "malformed face record", __FILE__, __LINE__);
            faces[3 * cf + c] = (int64_t)v - 1;
            q = next;
            while (*q && *q != ' ' && *q != '\t') ++q;      // skip "/vt/vn"
          }
        }
        ++cf;
      }
      *eol = saved;
    }
    p = eol + 1;
  }
  *nv = cv;
  *nf = cf;
  return 0;
}
