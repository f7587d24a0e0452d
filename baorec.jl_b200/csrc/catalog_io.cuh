// Catalog files either side of the path ("NPY/CSV readers, pinned-memory async I/O"): the internal
// C++ side of catalog_io.cu, shared with the file-driven batch pipeline in batch_files.cu.  Host code only.
#pragma once
#include <stddef.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "baorec_b200.h"

namespace baorec {
namespace io {

// Read-only memory map of a whole file.
struct MappedFile {
  const char* p = nullptr;
  size_t n = 0;
  int open(const char* path);  // BAOREC_OK / BAOREC_ERR_IO (message set)
  void close();
  ~MappedFile() { close(); }
  MappedFile() = default;
  MappedFile(const MappedFile&) = delete;
  MappedFile& operator=(const MappedFile&) = delete;
};

// threads actually used for `bytes` of work: n_threads <= 0 -> the host's cores; never more than one per 4 MiB
int resolve_threads(int n_threads, size_t bytes);

// ---- delimited text ---------------------------------------------------------------------------------
// Chunks of a text file that begin at line starts, and the data rows (non-blank lines not starting with '#') in each.
struct TextLayout {
  std::vector<size_t> begin;  // T + 1 offsets, begin[T] = file size
  std::vector<int64_t> rows;  // T counts
  int64_t n_rows = 0;
  int n_cols = 0;  // fields of the first data row
};
int text_layout(const MappedFile& f, char delim, int n_threads, TextLayout* out);
int text_read(const MappedFile& f, const TextLayout& lay, char delim, int n_out, const int* cols, float* const* out,
              int64_t capacity, const char* path);

// ---- NPY -----------------------------------------------------------------------------------------------
struct NpyHeader {
  int dtype = 0;  // BAOREC_DTYPE_F32 / BAOREC_DTYPE_F64
  int fortran_order = 0;
  int ndim = 0;
  int64_t shape[2] = {0, 0};
  size_t data_offset = 0;
  int64_t rows() const { return ndim == 0 ? 1 : shape[0]; }
  int64_t cols() const { return ndim == 2 ? shape[1] : 1; }
};
int npy_parse(const MappedFile& f, const char* path, NpyHeader* h);
int npy_read(const MappedFile& f, const NpyHeader& h, int n_out, const int* cols, float* const* out, int64_t capacity,
             int n_threads, const char* path);
int npy_write(const char* path, int n_cols, const float* const* cols, int64_t n, int n_threads);

// lo < key < hi, stable, in place; returns the rows kept
int64_t select_rows(int n_cols, float* const* cols, int64_t n, int key, float lo, float hi, int n_threads);

inline bool is_npy_path(const char* path) {
  const std::string s(path);
  return s.size() >= 4 && s.compare(s.size() - 4, 4, ".npy") == 0;
}

}  // namespace io
}  // namespace baorec
