// Catalog files either side of the path: what the reference's examples do with CSV.jl and NPZ.jl
// around run! / reconstructed_positions --
//   CSV.File(fn, delim = ' ', ignorerepeated = true, header = [...], types = [Float32 ...])   examples/simulation.jl:12-13,
//                                                                                               examples/lightcone.jl:22-23
//   data_cat[map(z -> ((z > 0.8) & (z < 1)), data_cat.z), :]                                    examples/lightcone.jl:25-26
//   npzwrite(fn, hcat(new_pos...))                                                              examples/simulation.jl:38-40
// as threaded host code that fills caller-owned (pinned) SoA Float32 arrays, i.e. the layout baorec_run_host_f32 /
// baorec_batch_host_f32 take.  No device code here: these entry points need neither a context nor a GPU.
#include "catalog_io.cuh"

#include <errno.h>
#include <stdlib.h>
#include <fcntl.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <charconv>
#include <thread>

#include "internal.cuh"

namespace baorec {
namespace io {

int MappedFile::open(const char* path) {
  close();
  const int fd = ::open(path, O_RDONLY);
  if (fd < 0) {
    set_error("cannot open '%s': %s", path, strerror(errno));
    return BAOREC_ERR_IO;
  }
  struct stat sb;
  if (fstat(fd, &sb) != 0 || !S_ISREG(sb.st_mode)) {
    set_error("'%s' is not a regular file", path);
    ::close(fd);
    return BAOREC_ERR_IO;
  }
  n = (size_t)sb.st_size;
  if (n > 0) {
    void* m = mmap(nullptr, n, PROT_READ, MAP_PRIVATE, fd, 0);
    if (m == MAP_FAILED) {
      set_error("mmap of '%s' (%zu bytes) failed: %s", path, n, strerror(errno));
      ::close(fd);
      n = 0;
      return BAOREC_ERR_IO;
    }
    madvise(m, n, MADV_SEQUENTIAL);
    p = static_cast<const char*>(m);
  }
  ::close(fd);
  return BAOREC_OK;
}

void MappedFile::close() {
  if (p) munmap(const_cast<char*>(p), n);
  p = nullptr;
  n = 0;
}

int resolve_threads(int n_threads, size_t bytes) {
  int t = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
  if (t < 1) t = 1;
  const size_t by_size = bytes / ((size_t)4 << 20) + 1;
  if ((size_t)t > by_size) t = (int)by_size;
  return t;
}

template <class F>
static void parallel_for(int T, F&& body) {
  if (T <= 1) {
    body(0);
    return;
  }
  std::vector<std::thread> th;
  th.reserve(T - 1);
  for (int t = 1; t < T; t++) th.emplace_back([&body, t] { body(t); });
  body(0);
  for (auto& x : th) x.join();
}

// ---- delimited text ------------------------------------------------------------------------------------
static inline bool is_blank(char c) { return c == ' ' || c == '\t' || c == '\r'; }

static inline bool is_data_line(const char* s, const char* e) {
  while (s < e && is_blank(*s)) s++;
  return s < e && *s != '#';
}

// Next field of a line.  delim == ' ': fields are runs of non-blank characters (CSV.jl's ignorerepeated = true).
// Any other delimiter separates exactly; blanks around a field are dropped.  Returns false at the end of the line.
static inline bool next_field(const char*& s, const char* e, char delim, const char*& fb, const char*& fe, bool& first) {
  if (delim == ' ') {
    while (s < e && is_blank(*s)) s++;
    if (s >= e) return false;
    fb = s;
    while (s < e && !is_blank(*s)) s++;
    fe = s;
    return true;
  }
  if (!first) {
    if (s >= e) return false;
    s++;  // the delimiter that ended the previous field
  } else if (s >= e) {
    return false;
  }
  first = false;
  const char* q = static_cast<const char*>(memchr(s, delim, (size_t)(e - s)));
  const char* end = q ? q : e;
  fb = s;
  fe = end;
  while (fb < fe && is_blank(*fb)) fb++;
  while (fe > fb && is_blank(fe[-1])) fe--;
  s = end;
  return true;
}

static int count_fields(const char* s, const char* e, char delim) {
  const char *fb, *fe;
  bool first = true;
  int k = 0;
  while (next_field(s, e, delim, fb, fe, first)) k++;
  return k;
}

int text_layout(const MappedFile& f, char delim, int n_threads, TextLayout* out) {
  const int T = resolve_threads(n_threads, f.n);
  out->begin.assign(T + 1, f.n);
  out->rows.assign(T, 0);
  out->begin[0] = 0;
  for (int t = 1; t < T; t++) {
    const size_t nominal = f.n / T * t;
    size_t b = f.n;
    if (nominal > 0 && nominal <= f.n) {
      const void* q = memchr(f.p + nominal - 1, '\n', f.n - (nominal - 1));
      if (q) b = (size_t)(static_cast<const char*>(q) - f.p) + 1;
    }
    out->begin[t] = std::max(b, out->begin[t - 1]);
  }
  std::vector<size_t> first_row(T, f.n);
  parallel_for(T, [&](int t) {
    const char* s = f.p + out->begin[t];
    const char* const e = f.p + out->begin[t + 1];
    int64_t r = 0;
    while (s < e) {
      const char* nl = static_cast<const char*>(memchr(s, '\n', (size_t)(e - s)));
      const char* le = nl ? nl : e;
      if (is_data_line(s, le)) {
        if (r == 0) first_row[t] = (size_t)(s - f.p);
        r++;
      }
      s = le + 1;
    }
    out->rows[t] = r;
  });
  out->n_rows = 0;
  out->n_cols = 0;
  for (int t = 0; t < T; t++) out->n_rows += out->rows[t];
  for (int t = 0; t < T; t++)
    if (out->rows[t] > 0) {
      const char* s = f.p + first_row[t];
      const char* nl = static_cast<const char*>(memchr(s, '\n', f.n - first_row[t]));
      out->n_cols = count_fields(s, nl ? nl : f.p + f.n, delim);
      break;
    }
  return BAOREC_OK;
}

int text_read(const MappedFile& f, const TextLayout& lay, char delim, int n_out, const int* cols, float* const* out,
              int64_t capacity, const char* path) {
  if (lay.n_rows > capacity) {
    set_error("'%s' holds %lld rows, the arrays passed hold %lld", path, (long long)lay.n_rows, (long long)capacity);
    return BAOREC_ERR_INVALID;
  }
  const int T = (int)lay.rows.size();
  int maxcol = 0;
  for (int j = 0; j < n_out; j++) maxcol = std::max(maxcol, cols[j]);
  std::vector<int64_t> row0(T, 0);
  for (int t = 1; t < T; t++) row0[t] = row0[t - 1] + lay.rows[t - 1];
  // first failure per thread: byte offset of the line and what was wrong with it
  std::vector<size_t> bad_at(T, (size_t)-1);
  std::vector<int> bad_field(T, -1);
  parallel_for(T, [&](int t) {
    const char* s = f.p + lay.begin[t];
    const char* const e = f.p + lay.begin[t + 1];
    int64_t r = row0[t];
    while (s < e) {
      const char* nl = static_cast<const char*>(memchr(s, '\n', (size_t)(e - s)));
      const char* le = nl ? nl : e;
      if (is_data_line(s, le)) {
        const char *q = s, *fb, *fe;
        bool first = true;
        int k = 0;
        for (; k <= maxcol; k++) {
          if (!next_field(q, le, delim, fb, fe, first)) break;
          bool wanted = false;
          for (int j = 0; j < n_out; j++) wanted |= cols[j] == k;
          if (!wanted) continue;
          if (fb < fe && *fb == '+') fb++;
          float v;
          const auto res = std::from_chars(fb, fe, v);
          if (res.ec == std::errc::result_out_of_range && res.ptr == fe) {
            // from_chars leaves v untouched on overflow / underflow: take strtod's answer (inf or a denormal)
            v = (float)strtod(std::string(fb, fe).c_str(), nullptr);
          } else if (res.ec != std::errc() || res.ptr != fe) {
            break;
          }
          for (int j = 0; j < n_out; j++)
            if (cols[j] == k) out[j][r] = v;
        }
        if (k <= maxcol) {
          bad_at[t] = (size_t)(s - f.p);
          bad_field[t] = k;
          return;
        }
        r++;
      }
      s = le + 1;
    }
  });
  for (int t = 0; t < T; t++)
    if (bad_at[t] != (size_t)-1) {
      int64_t line = 1;
      for (const char* c = f.p; c < f.p + bad_at[t]; c++) line += *c == '\n';
      const char* s = f.p + bad_at[t];
      const char* nl = static_cast<const char*>(memchr(s, '\n', f.n - bad_at[t]));
      const size_t len = std::min<size_t>((size_t)((nl ? nl : f.p + f.n) - s), 80);
      set_error("'%s' line %lld: field %d is missing or not a number: \"%.*s\"", path, (long long)line, bad_field[t] + 1,
                (int)len, s);
      return BAOREC_ERR_IO;
    }
  return BAOREC_OK;
}

// ---- NPY (format 1.0 / 2.0 / 3.0) ----------------------------------------------------------------------------
static bool find_value(const std::string& h, const char* key, size_t* pos) {
  size_t k = h.find(key);
  if (k == std::string::npos) return false;
  k = h.find(':', k);
  if (k == std::string::npos) return false;
  k++;
  while (k < h.size() && h[k] == ' ') k++;
  *pos = k;
  return true;
}

int npy_parse(const MappedFile& f, const char* path, NpyHeader* h) {
  static const unsigned char magic[6] = {0x93, 'N', 'U', 'M', 'P', 'Y'};
  if (f.n < 10 || memcmp(f.p, magic, 6) != 0) {
    set_error("'%s' is not an NPY file", path);
    return BAOREC_ERR_IO;
  }
  const unsigned char* u = reinterpret_cast<const unsigned char*>(f.p);
  const int major = u[6];
  size_t hlen, hoff;
  if (major == 1) {
    hlen = (size_t)u[8] | ((size_t)u[9] << 8);
    hoff = 10;
  } else if (major == 2 || major == 3) {
    if (f.n < 12) {
      set_error("'%s': truncated NPY header", path);
      return BAOREC_ERR_IO;
    }
    hlen = (size_t)u[8] | ((size_t)u[9] << 8) | ((size_t)u[10] << 16) | ((size_t)u[11] << 24);
    hoff = 12;
  } else {
    set_error("'%s': NPY format version %d is not supported", path, major);
    return BAOREC_ERR_IO;
  }
  if (hoff + hlen > f.n) {
    set_error("'%s': truncated NPY header", path);
    return BAOREC_ERR_IO;
  }
  const std::string hd(f.p + hoff, hlen);
  size_t k;
  if (!find_value(hd, "'descr'", &k) || k + 4 >= hd.size() || (hd[k] != '\'' && hd[k] != '"')) {
    set_error("'%s': NPY header without a simple 'descr'", path);
    return BAOREC_ERR_IO;
  }
  const std::string descr = hd.substr(k + 1, 3);
  if (descr == "<f4" || descr == "=f4" || descr == "|f4") h->dtype = BAOREC_DTYPE_F32;
  else if (descr == "<f8" || descr == "=f8" || descr == "|f8") h->dtype = BAOREC_DTYPE_F64;
  else {
    set_error("'%s': NPY dtype '%s' is not supported (little-endian f4 / f8 only)", path, descr.c_str());
    return BAOREC_ERR_IO;
  }
  if (!find_value(hd, "'fortran_order'", &k)) {
    set_error("'%s': NPY header without 'fortran_order'", path);
    return BAOREC_ERR_IO;
  }
  h->fortran_order = hd.compare(k, 4, "True") == 0;
  if (!find_value(hd, "'shape'", &k) || hd[k] != '(') {
    set_error("'%s': NPY header without 'shape'", path);
    return BAOREC_ERR_IO;
  }
  const size_t close = hd.find(')', k);
  if (close == std::string::npos) {
    set_error("'%s': NPY header with an unterminated shape", path);
    return BAOREC_ERR_IO;
  }
  h->ndim = 0;
  h->shape[0] = h->shape[1] = 0;
  size_t q = k + 1;
  while (q < close) {
    while (q < close && (hd[q] == ' ' || hd[q] == ',')) q++;
    if (q >= close) break;
    long long v = 0;
    const auto res = std::from_chars(hd.data() + q, hd.data() + close, v);
    if (res.ec != std::errc() || v < 0) {
      set_error("'%s': cannot read the NPY shape", path);
      return BAOREC_ERR_IO;
    }
    // numpy writes "3L" on very old versions
    q = (size_t)(res.ptr - hd.data());
    if (q < close && hd[q] == 'L') q++;
    if (h->ndim >= 2) {
      set_error("'%s': NPY arrays of more than two dimensions are not catalogs", path);
      return BAOREC_ERR_IO;
    }
    h->shape[h->ndim++] = v;
  }
  h->data_offset = hoff + hlen;
  // rows * cols * itemsize <= bytes present, without trusting the product not to overflow
  const size_t avail = f.n - h->data_offset, r = (size_t)h->rows(), c = (size_t)h->cols();
  if (r != 0 && c != 0 && (c > avail / (size_t)h->dtype || r > avail / (size_t)h->dtype / c)) {
    set_error("'%s': NPY data truncated (shape (%zu, %zu) of %d-byte items, %zu bytes present)", path, r, c, h->dtype, avail);
    return BAOREC_ERR_IO;
  }
  return BAOREC_OK;
}

template <class S>
static void convert_block(const char* base, size_t first, size_t stride, int64_t r0, int64_t r1, float* dst) {
  // element r of the column sits at base + (first + r * stride) * sizeof(S); unaligned reads go through memcpy
  if (stride == 1 && sizeof(S) == sizeof(float)) {
    memcpy(dst + r0, base + (first + (size_t)r0) * sizeof(S), (size_t)(r1 - r0) * sizeof(float));
    return;
  }
  for (int64_t r = r0; r < r1; r++) {
    S v;
    memcpy(&v, base + (first + (size_t)r * stride) * sizeof(S), sizeof(S));
    dst[r] = (float)v;
  }
}

int npy_read(const MappedFile& f, const NpyHeader& h, int n_out, const int* cols, float* const* out, int64_t capacity,
             int n_threads, const char* path) {
  const int64_t N = h.rows(), K = h.cols();
  if (N > capacity) {
    set_error("'%s' holds %lld rows, the arrays passed hold %lld", path, (long long)N, (long long)capacity);
    return BAOREC_ERR_INVALID;
  }
  for (int j = 0; j < n_out; j++)
    if (cols[j] < 0 || cols[j] >= K) {
      set_error("'%s' has %lld columns, column %d requested", path, (long long)K, cols[j]);
      return BAOREC_ERR_INVALID;
    }
  const char* base = f.p + h.data_offset;
  const int T = resolve_threads(n_threads, (size_t)N * (size_t)n_out * (size_t)h.dtype);
  if (h.dtype == BAOREC_DTYPE_F32 && (h.fortran_order || K == 1) && N > 0) {
    // the columns lie in the file as they will lie in memory: pread straight into the (pinned) arrays, one row range
    // per thread -- no page of the file is mapped (mmap + memcpy measured 1.3 GB/s on the first hardware run: one
    // minor fault per 4 KiB)
    const int fd = ::open(path, O_RDONLY);
    if (fd >= 0) {
      std::vector<int> bad(T, 0);
      parallel_for(T, [&](int t) {
        const int64_t r0 = N * t / T, r1 = N * (t + 1) / T;
        for (int j = 0; j < n_out && !bad[t]; j++) {
          char* dst = reinterpret_cast<char*>(out[j] + r0);
          size_t left = (size_t)(r1 - r0) * sizeof(float);
          off_t off = (off_t)(h.data_offset + ((size_t)cols[j] * (size_t)N + (size_t)r0) * sizeof(float));
          while (left > 0) {
            const ssize_t got = pread(fd, dst, std::min<size_t>(left, (size_t)1 << 30), off);
            if (got < 0 && errno == EINTR) continue;
            if (got <= 0) {
              bad[t] = 1;
              break;
            }
            dst += got;
            off += got;
            left -= (size_t)got;
          }
        }
      });
      ::close(fd);
      bool ok = true;
      for (int t = 0; t < T; t++) ok &= !bad[t];
      if (ok) return BAOREC_OK;
      set_error("reading '%s' failed: %s", path, strerror(errno));
      return BAOREC_ERR_IO;
    }
  }
  parallel_for(T, [&](int t) {
    const int64_t r0 = N * t / T, r1 = N * (t + 1) / T;
    for (int j = 0; j < n_out; j++) {
      const size_t first = h.fortran_order ? (size_t)cols[j] * (size_t)N : (size_t)cols[j];
      const size_t stride = h.fortran_order ? 1 : (size_t)K;
      if (h.dtype == BAOREC_DTYPE_F32) convert_block<float>(base, first, stride, r0, r1, out[j]);
      else convert_block<double>(base, first, stride, r0, r1, out[j]);
    }
  });
  return BAOREC_OK;
}

static int write_all(int fd, const void* buf, size_t n, const char* path) {
  const char* c = static_cast<const char*>(buf);
  while (n > 0) {
    const ssize_t w = ::write(fd, c, std::min<size_t>(n, (size_t)1 << 30));
    if (w < 0) {
      if (errno == EINTR) continue;
      set_error("writing '%s' failed: %s", path, strerror(errno));
      return BAOREC_ERR_IO;
    }
    c += w;
    n -= (size_t)w;
  }
  return BAOREC_OK;
}

// npzwrite(fn, hcat(cols...)): an (n, n_cols) Float32 matrix in Julia's column-major memory, which NPZ.jl stores as it
// lies with 'fortran_order': True -- i.e. the SoA columns one after the other.  One column is written as a plain vector.
int npy_write(const char* path, int n_cols, const float* const* cols, int64_t n, int n_threads) {
  char dict[160];
  if (n_cols == 1) snprintf(dict, sizeof dict, "{'descr': '<f4', 'fortran_order': False, 'shape': (%lld,), }", (long long)n);
  else snprintf(dict, sizeof dict, "{'descr': '<f4', 'fortran_order': True, 'shape': (%lld, %d), }", (long long)n, n_cols);
  std::string hd(dict);
  const size_t total = (10 + hd.size() + 1 + 63) / 64 * 64;  // magic + version + length, dictionary, '\n'
  hd.append(total - 10 - hd.size() - 1, ' ');
  hd.push_back('\n');
  unsigned char pre[10] = {0x93, 'N', 'U', 'M', 'P', 'Y', 1, 0, (unsigned char)(hd.size() & 0xff), (unsigned char)(hd.size() >> 8)};
  const int fd = ::open(path, O_WRONLY | O_CREAT | O_TRUNC, 0644);
  if (fd < 0) {
    set_error("cannot create '%s': %s", path, strerror(errno));
    return BAOREC_ERR_IO;
  }
  int s = write_all(fd, pre, sizeof pre, path);
  if (s == BAOREC_OK) s = write_all(fd, hd.data(), hd.size(), path);
  // the columns: one row range per thread, each at its place in the file
  const size_t col_bytes = (size_t)n * sizeof(float);
  const int T = resolve_threads(n_threads, col_bytes * (size_t)n_cols);
  std::vector<int> err(T, 0);
  if (s == BAOREC_OK && n > 0)
    parallel_for(T, [&](int t) {
      const int64_t r0 = n * t / T, r1 = n * (t + 1) / T;
      for (int c = 0; c < n_cols && !err[t]; c++) {
        const char* src = reinterpret_cast<const char*>(cols[c] + r0);
        size_t left = (size_t)(r1 - r0) * sizeof(float);
        off_t off = (off_t)(10 + hd.size() + (size_t)c * col_bytes + (size_t)r0 * sizeof(float));
        while (left > 0) {
          const ssize_t w = pwrite(fd, src, std::min<size_t>(left, (size_t)1 << 30), off);
          if (w < 0 && errno == EINTR) continue;
          if (w <= 0) {
            err[t] = errno ? errno : EIO;
            break;
          }
          src += w;
          off += w;
          left -= (size_t)w;
        }
      }
    });
  for (int t = 0; t < T && s == BAOREC_OK; t++)
    if (err[t]) {
      set_error("writing '%s' failed: %s", path, strerror(err[t]));
      s = BAOREC_ERR_IO;
    }
  if (::close(fd) != 0 && s == BAOREC_OK) {
    set_error("closing '%s' failed: %s", path, strerror(errno));
    s = BAOREC_ERR_IO;
  }
  return s;
}

// ---- row selection ---------------------------------------------------------------------------------------
int64_t select_rows(int n_cols, float* const* cols, int64_t n, int key, float lo, float hi, int n_threads) {
  const int T = resolve_threads(n_threads, (size_t)n * (size_t)n_cols * sizeof(float));
  std::vector<int64_t> kept(T, 0);
  // every thread compacts its own block of rows in place (the key column last: the others are tested against it) ...
  parallel_for(T, [&](int t) {
    const int64_t r0 = n * t / T, r1 = n * (t + 1) / T;
    const float* k = cols[key];
    int64_t m = 0;
    for (int c = 0; c < n_cols; c++) {
      if (c == key) continue;
      float* a = cols[c];
      m = r0;
      for (int64_t r = r0; r < r1; r++)
        if (k[r] > lo && k[r] < hi) a[m++] = a[r];
    }
    float* a = cols[key];
    m = r0;
    for (int64_t r = r0; r < r1; r++) {
      const float v = a[r];
      if (v > lo && v < hi) a[m++] = v;
    }
    kept[t] = m - r0;
  });
  // ... then the blocks close up, left to right
  int64_t total = kept[0];
  for (int t = 1; t < T; t++) {
    const int64_t r0 = n * t / T;
    if (kept[t] > 0 && r0 != total)
      for (int c = 0; c < n_cols; c++) memmove(cols[c] + total, cols[c] + r0, (size_t)kept[t] * sizeof(float));
    total += kept[t];
  }
  return total;
}

}  // namespace io
}  // namespace baorec

using namespace baorec;

extern "C" {

int baorec_text_catalog_scan(const char* path, char delim, int64_t* n_rows, int* n_cols, int n_threads) {
  BR_REQUIRE(path && n_rows, "path / n_rows");
  io::MappedFile f;
  BR_TRY(f.open(path));
  io::TextLayout lay;
  BR_TRY(io::text_layout(f, delim, n_threads, &lay));
  *n_rows = lay.n_rows;
  if (n_cols) *n_cols = lay.n_cols;
  return BAOREC_OK;
}

int baorec_text_catalog_read_f32(const char* path, char delim, int n_out, const int* cols, float* const* h_out,
                                 int64_t capacity, int64_t* n_read, int n_threads) {
  BR_REQUIRE(path && cols && h_out && n_read, "path / cols / h_out / n_read");
  BR_REQUIRE(n_out >= 1 && n_out <= 64, "1 <= n_out <= 64");
  for (int j = 0; j < n_out; j++) BR_REQUIRE(cols[j] >= 0 && (h_out[j] || capacity == 0), "column index / output array");
  io::MappedFile f;
  BR_TRY(f.open(path));
  io::TextLayout lay;
  BR_TRY(io::text_layout(f, delim, n_threads, &lay));
  BR_TRY(io::text_read(f, lay, delim, n_out, cols, h_out, capacity, path));
  *n_read = lay.n_rows;
  return BAOREC_OK;
}

int baorec_npy_info(const char* path, int* dtype, int* fortran_order, int64_t* n_rows, int64_t* n_cols) {
  BR_REQUIRE(path != nullptr, "path");
  io::MappedFile f;
  BR_TRY(f.open(path));
  io::NpyHeader h;
  BR_TRY(io::npy_parse(f, path, &h));
  if (dtype) *dtype = h.dtype;
  if (fortran_order) *fortran_order = h.fortran_order;
  if (n_rows) *n_rows = h.rows();
  if (n_cols) *n_cols = h.cols();
  return BAOREC_OK;
}

int baorec_npy_read_columns_f32(const char* path, int n_out, const int* cols, float* const* h_out, int64_t capacity,
                                int64_t* n_read, int n_threads) {
  BR_REQUIRE(path && cols && h_out && n_read, "path / cols / h_out / n_read");
  BR_REQUIRE(n_out >= 1 && n_out <= 64, "1 <= n_out <= 64");
  for (int j = 0; j < n_out; j++) BR_REQUIRE(h_out[j] != nullptr || capacity == 0, "output array");
  io::MappedFile f;
  BR_TRY(f.open(path));
  io::NpyHeader h;
  BR_TRY(io::npy_parse(f, path, &h));
  BR_TRY(io::npy_read(f, h, n_out, cols, h_out, capacity, n_threads, path));
  *n_read = h.rows();
  return BAOREC_OK;
}

int baorec_npy_write_columns_f32(const char* path, int n_cols, const float* const* h_cols, int64_t n) {
  BR_REQUIRE(path && h_cols, "path / h_cols");
  BR_REQUIRE(n_cols >= 1 && n_cols <= 64 && n >= 0, "1 <= n_cols <= 64, n >= 0");
  for (int j = 0; j < n_cols; j++) BR_REQUIRE(h_cols[j] != nullptr || n == 0, "column array");
  return io::npy_write(path, n_cols, h_cols, n, 0);
}

int baorec_catalog_select_f32(int n_cols, float* const* h_cols, int64_t n, int key, float lo, float hi, int64_t* n_kept,
                              int n_threads) {
  BR_REQUIRE(h_cols && n_kept, "h_cols / n_kept");
  BR_REQUIRE(n_cols >= 1 && n_cols <= 64 && n >= 0 && key >= 0 && key < n_cols, "1 <= n_cols <= 64, 0 <= key < n_cols");
  for (int j = 0; j < n_cols; j++) BR_REQUIRE(h_cols[j] != nullptr || n == 0, "column array");
  *n_kept = n > 0 ? io::select_rows(n_cols, h_cols, n, key, lo, hi, n_threads) : 0;
  return BAOREC_OK;
}

}  // extern "C"
