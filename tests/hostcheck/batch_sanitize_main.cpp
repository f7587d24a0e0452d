// Driver for the sanitizer runs of the file-driven batch source (tests/test_batch_files_host.py::test_io_threads_under_sanitizers):
// the host check of batch_hostcheck.cpp called over and over -- good batches with 3, 4, 5 buffer sets and 1 - 4 parser
// threads, a failing device step, a bad line and a missing file in the middle -- in a binary built with
// -fsanitize=thread or -fsanitize=address,undefined.  argv[1] = directory prepared by the test (m0 ... m7, bad.dat).
#include <stdio.h>
#include <stdint.h>
#include <string>
#include <vector>
extern "C" int hc_batch_files(int n_catalogs, const char* const* in_paths, char delim, const int* cols,
                              const char* const* out_paths, int n_slots, int n_threads, int fail_at, int64_t* n_rows, double* seconds);
extern "C" const char* baorec_last_error(void);
int main(int argc, char** argv) {
  const std::string dir = argc > 1 ? argv[1] : "/tmp/san";
  std::vector<std::string> in, out;
  const char* ext[8] = {"dat","dat","npy","dat","dat","npy","dat","dat"};
  for (int i = 0; i < 8; i++) { in.push_back(dir + "/m" + std::to_string(i) + "." + ext[i]); out.push_back(dir + "/o" + std::to_string(i) + ".npy"); }
  int cols[4] = {0, 1, 2, 3};
  int bad = 0;
  for (int rep = 0; rep < 6; rep++) {
    for (int slots = 3; slots <= 5; slots++) {
      std::vector<const char*> pi, po;
      for (auto& s : in) pi.push_back(s.c_str());
      for (auto& s : out) po.push_back(s.c_str());
      int64_t rows[8]; double sec[4];
      int rc = hc_batch_files(8, pi.data(), ' ', cols, po.data(), slots, 1 + rep % 4, -1, rows, sec);
      if (rc) { printf("rc=%d %s\n", rc, baorec_last_error()); bad++; }
      // failing device step, and a bad file in the middle
      rc = hc_batch_files(8, pi.data(), ' ', cols, po.data(), slots, 2, 3, rows, sec);
      if (rc != -5) { printf("expected -5, got %d\n", rc); bad++; }
      const std::string badf = dir + "/bad.dat", nonef = dir + "/none.dat";
      pi[4] = badf.c_str();
      rc = hc_batch_files(8, pi.data(), ' ', cols, po.data(), slots, 2, -1, rows, sec);
      if (rc != -9) { printf("expected -9, got %d\n", rc); bad++; }
      pi[4] = nonef.c_str();
      rc = hc_batch_files(8, pi.data(), ' ', cols, nullptr, slots, 2, -1, rows, sec);
      if (rc != -9) { printf("expected -9, got %d\n", rc); bad++; }
    }
  }
  printf("done bad=%d\n", bad);
  return bad;
}
