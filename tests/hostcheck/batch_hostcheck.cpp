// Host check of the file-driven batch source (baorec.jl_b200/csrc/batch.cuh: FileBatchSource, the reader / writer
// threads behind baorec_batch_files_f32) on a machine without a GPU: the very class the library uses, with malloc in
// place of pinned memory, driven in the call order of batch_pipeline (api.cu) -- acquire(0); per catalog: acquire(i+1),
// "reconstruct" catalog i, release(i-1); release(last) -- by a stand-in that writes pos + w * (1, 2, 3) as the result.
// Compiled together with csrc/catalog_io.cu (plain C++: the file holds no device code) by tests/test_batch_files_host.py.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>

#include "batch.cuh"

namespace baorec {
static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
}
}  // namespace baorec

extern "C" const char* baorec_last_error(void) { return baorec::g_err; }

static int hc_alloc(void** out, size_t bytes) {
  *out = malloc(bytes);
  return *out ? 0 : 1;
}

// fail_at >= 0: the stand-in "device" fails at that catalog (what an out-of-box particle does to the real pipeline).
extern "C" int hc_batch_files(int n_catalogs, const char* const* in_paths, char delim, const int* cols,
                              const char* const* out_paths, int n_slots, int n_threads, int fail_at, int64_t* n_rows,
                              double* seconds) {
  using namespace baorec;
  FileBatchConfig cfg;
  cfg.n_catalogs = n_catalogs;
  cfg.in_paths = in_paths;
  cfg.out_paths = out_paths;
  cfg.delim = delim;
  for (int c = 0; c < 4; c++) cfg.cols[c] = cols[c];
  cfg.n_slots = n_slots;
  cfg.n_threads = n_threads;
  HostAllocator al{hc_alloc, free};
  FileBatchSource src(cfg, al);
  BatchItem item[2];
  auto body = [&]() -> int {
    int s = src.acquire(0, &item[0]);
    if (s != BAOREC_OK) return s;
    for (int i = 0; i < n_catalogs; i++) {
      const int b = i & 1;
      if (i + 1 < n_catalogs && (s = src.acquire(i + 1, &item[b ^ 1])) != BAOREC_OK) return s;
      if (i == fail_at) {
        set_error("stand-in failure at catalog %d", i);
        return BAOREC_ERR_OUT_OF_BOX;
      }
      const BatchItem& it = item[b];
      for (int64_t r = 0; r < it.n; r++) {
        it.ox[r] = it.x[r] + it.w[r] * 1.0f;
        it.oy[r] = it.y[r] + it.w[r] * 2.0f;
        it.oz[r] = it.z[r] + it.w[r] * 3.0f;
      }
      if (i >= 1 && (s = src.release(i - 1)) != BAOREC_OK) return s;
    }
    return src.release(n_catalogs - 1);
  };
  int status = body();
  if (status == BAOREC_OK) status = src.finish();
  else {
    const std::string msg = baorec_last_error();
    src.finish(true);
    set_error("%s", msg.c_str());
  }
  if (n_rows)
    for (int i = 0; i < n_catalogs; i++) n_rows[i] = src.rows()[i];
  if (seconds) {
    seconds[0] = src.read_seconds();
    seconds[1] = src.write_seconds();
    seconds[2] = src.wait_seconds();
  }
  return status;
}
