// Many catalogs per process (README.md:11 of the reference: "one process, many reconstructions"): the software-pipelined
// host pipeline of api.cu (batch_pipeline) takes its catalogs from a BatchSource -- host arrays the caller already holds
// (baorec_batch_host_f32) or catalog FILES read, parsed and written by I/O threads while the device works on their
// neighbours (baorec_batch_files_f32; SURVEY.md 8f N4 "many-mock batch driver with pinned-memory async I/O").
//
// FileBatchSource is host code without any CUDA call of its own (pinned memory comes from the allocator it is given),
// so that tests/hostcheck/batch_hostcheck.cpp can drive the very same class on a machine without a GPU.
#pragma once
#include <stdint.h>
#include <string.h>

#include <chrono>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "baorec_b200.h"
#include "catalog_io.cuh"
#include "host_slot.cuh"

struct baorec_ctx;

namespace baorec {

void set_error(const char* fmt, ...);

// One catalog as the pipeline sees it: SoA host arrays of n floats (pinned for the copies to be asynchronous).
struct BatchItem {
  float *x = nullptr, *y = nullptr, *z = nullptr;
  const float* w = nullptr;
  int64_t n = 0;
  float *ox = nullptr, *oy = nullptr, *oz = nullptr;
};

struct BatchSource {
  virtual ~BatchSource() {}
  // Catalog i is in host memory (blocks until it is).  Called once per catalog, in order, up to one catalog ahead of
  // the one being reconstructed.
  virtual int acquire(int i, BatchItem* it) = 0;
  // The results of catalog i (and its wrapped positions, if cic! wrapped any) are complete in host memory; the
  // pipeline does not touch the arrays of catalog i again.  Called once per acquired catalog, in order.
  virtual int release(int i) = 0;
  // Largest catalog if known in advance (device staging is then allocated once), else 0 (grown when needed).
  virtual int64_t max_rows() const { return 0; }
};

// api.cu
int batch_pipeline(baorec_ctx* ctx, const baorec_params* p, int algorithm, int n_catalogs, int field, int shifts_only,
                   BatchSource& src);

struct HostAllocator {
  int (*alloc)(void** out, size_t bytes);  // 0 = ok
  void (*free)(void* p);
};

struct FileBatchConfig {
  int n_catalogs = 0;
  const char* const* in_paths = nullptr;
  const char* const* out_paths = nullptr;  // NULL, or per catalog NULL = nothing written
  char delim = ' ';
  int cols[4] = {0, 1, 2, -1};  // file columns of x, y, z, w; w < 0: weights of one (examples/simulation.jl:16)
  int n_slots = 4;              // host buffer sets: reading ahead | uploading | in flight | being written
  int n_threads = 0;            // parser threads of the reader
};

// Reader thread: file i -> slot (7 arrays x, y, z, w, ox, oy, oz of `cap` floats in one allocation).  Main thread:
// acquire / release.  Writer thread: released slot -> NPY file, slot back to the reader.
class FileBatchSource : public BatchSource {
 public:
  FileBatchSource(const FileBatchConfig& cfg, HostAllocator al, std::vector<HostSlot>* keep = nullptr)
      : cfg_(cfg), al_(al), keep_(keep) {
    if (cfg_.n_slots < 3) cfg_.n_slots = 3;  // acquire(i + 1) happens while i is in flight and i - 1 awaits its release
    slots_.resize(cfg_.n_slots);
    if (keep_) {  // adopt the buffer sets of an earlier call
      for (size_t s = 0; s < keep_->size(); s++) {
        if (s < slots_.size()) {
          slots_[s].base = (*keep_)[s].base;
          slots_[s].cap = (*keep_)[s].cap;
        } else if ((*keep_)[s].base) {
          al_.free((*keep_)[s].base);
        }
      }
      keep_->clear();
    }
    slot_of_.assign(cfg_.n_catalogs, -1);
    rows_.assign(cfg_.n_catalogs, 0);
    for (int s = 0; s < cfg_.n_slots; s++) free_.push_back(s);
    reader_ = std::thread([this] { read_loop(); });
    writer_ = std::thread([this] { write_loop(); });
  }
  ~FileBatchSource() override {
    finish(true);
    for (auto& s : slots_) {
      if (keep_) keep_->push_back(HostSlot{s.base, s.cap});
      else if (s.base) al_.free(s.base);
    }
  }

  int acquire(int i, BatchItem* it) override {
    const auto t0 = clock::now();
    std::unique_lock<std::mutex> lk(m_);
    cv_.wait(lk, [&] { return slot_of_[i] >= 0 || status_ != BAOREC_OK; });
    wait_s_ += seconds_since(t0);
    if (slot_of_[i] < 0) return fail_locked();
    const Slot& s = slots_[slot_of_[i]];
    float* b = s.base;
    it->x = b;
    it->y = b + s.cap;
    it->z = b + 2 * s.cap;
    it->w = b + 3 * s.cap;
    it->ox = b + 4 * s.cap;
    it->oy = b + 5 * s.cap;
    it->oz = b + 6 * s.cap;
    it->n = s.n;
    return BAOREC_OK;
  }

  int release(int i) override {
    std::lock_guard<std::mutex> lk(m_);
    to_write_.push_back(i);
    cv_.notify_all();
    return BAOREC_OK;
  }

  // Waits for the files of every released catalog to be written (abort: drops what is queued), joins the threads and
  // returns the first I/O error, message set on the calling thread.
  int finish(bool abort = false) {
    {
      std::lock_guard<std::mutex> lk(m_);
      if (abort) aborted_ = true;
      closing_ = true;
      cv_.notify_all();
    }
    if (reader_.joinable()) reader_.join();
    if (writer_.joinable()) writer_.join();
    std::lock_guard<std::mutex> lk(m_);
    return status_ == BAOREC_OK ? BAOREC_OK : fail_locked();
  }

  const std::vector<int64_t>& rows() const { return rows_; }
  double read_seconds() const { return read_s_; }
  double write_seconds() const { return write_s_; }
  double wait_seconds() const { return wait_s_; }

 private:
  using clock = std::chrono::steady_clock;
  static double seconds_since(clock::time_point t0) { return std::chrono::duration<double>(clock::now() - t0).count(); }

  struct Slot {
    float* base = nullptr;
    int64_t cap = 0, n = 0;
  };

  int fail_locked() {
    set_error("%s", message_.c_str());
    return status_;
  }
  void record_failure(int status) {  // on the I/O thread that saw it: its thread-local message travels with the status
    std::lock_guard<std::mutex> lk(m_);
    if (status_ == BAOREC_OK) {
      status_ = status;
      message_ = baorec_last_error();
    }
    cv_.notify_all();
  }

  int read_into(int i, Slot& s) {
    const char* path = cfg_.in_paths[i];
    io::MappedFile f;
    int st = f.open(path);
    if (st != BAOREC_OK) return st;
    const bool npy = io::is_npy_path(path);
    io::NpyHeader h;
    io::TextLayout lay;
    int64_t n;
    if (npy) {
      if ((st = io::npy_parse(f, path, &h)) != BAOREC_OK) return st;
      n = h.rows();
    } else {
      if ((st = io::text_layout(f, cfg_.delim, cfg_.n_threads, &lay)) != BAOREC_OK) return st;
      n = lay.n_rows;
    }
    if (n <= 0) {
      set_error("'%s' holds no rows", path);
      return BAOREC_ERR_IO;
    }
    if (n > s.cap) {  // the slot is free: nothing in flight touches it
      if (s.base) al_.free(s.base);
      s.base = nullptr;
      s.cap = 0;
      const int64_t cap = (n + n / 16 + 63) / 64 * 64;  // some head room: mocks of one suite differ by a few per cent
      void* mem = nullptr;
      if (al_.alloc(&mem, (size_t)cap * 7 * sizeof(float)) != 0 || !mem) {
        set_error("host allocation of %lld bytes for a catalog of %lld rows failed", (long long)cap * 28, (long long)n);
        return BAOREC_ERR_NOMEM;
      }
      s.base = static_cast<float*>(mem);
      s.cap = cap;
    }
    s.n = n;
    const int nw = cfg_.cols[3] >= 0 ? 4 : 3;
    float* out[4] = {s.base, s.base + s.cap, s.base + 2 * s.cap, s.base + 3 * s.cap};
    st = npy ? io::npy_read(f, h, nw, cfg_.cols, out, s.cap, cfg_.n_threads, path)
             : io::text_read(f, lay, cfg_.delim, nw, cfg_.cols, out, s.cap, path);
    if (st != BAOREC_OK) return st;
    if (nw == 3)
      for (int64_t r = 0; r < n; r++) out[3][r] = 1.0f;
    return BAOREC_OK;
  }

  void read_loop() {
    for (int i = 0; i < cfg_.n_catalogs; i++) {
      int slot;
      {
        std::unique_lock<std::mutex> lk(m_);
        cv_.wait(lk, [&] { return !free_.empty() || closing_ || status_ != BAOREC_OK; });
        if (closing_ || status_ != BAOREC_OK) return;
        slot = free_.front();
        free_.pop_front();
      }
      const auto t0 = clock::now();
      const int st = read_into(i, slots_[slot]);
      if (st != BAOREC_OK) {
        record_failure(st);
        return;
      }
      std::lock_guard<std::mutex> lk(m_);
      read_s_ += seconds_since(t0);
      rows_[i] = slots_[slot].n;
      slot_of_[i] = slot;
      cv_.notify_all();
    }
  }

  void write_loop() {
    int written = 0;
    while (written < cfg_.n_catalogs) {
      int i;
      {
        std::unique_lock<std::mutex> lk(m_);
        cv_.wait(lk, [&] { return !to_write_.empty() || closing_ || status_ != BAOREC_OK; });
        if (to_write_.empty() || aborted_ || status_ != BAOREC_OK) return;  // closing with an empty queue: done
        i = to_write_.front();
        to_write_.pop_front();
      }
      const Slot& s = slots_[slot_of_[i]];
      const char* path = cfg_.out_paths ? cfg_.out_paths[i] : nullptr;
      if (path) {
        const auto t0 = clock::now();
        const float* cols[3] = {s.base + 4 * s.cap, s.base + 5 * s.cap, s.base + 6 * s.cap};
        const int st = io::npy_write(path, 3, cols, s.n, cfg_.n_threads);
        if (st != BAOREC_OK) {
          record_failure(st);
          return;
        }
        std::lock_guard<std::mutex> lk(m_);
        write_s_ += seconds_since(t0);
      }
      written++;
      std::lock_guard<std::mutex> lk(m_);
      free_.push_back(slot_of_[i]);
      cv_.notify_all();
    }
  }

  FileBatchConfig cfg_;
  HostAllocator al_;
  std::vector<HostSlot>* keep_;
  std::vector<Slot> slots_;
  std::vector<int> slot_of_;  // catalog -> slot once read
  std::vector<int64_t> rows_;
  std::deque<int> free_, to_write_;
  std::mutex m_;
  std::condition_variable cv_;
  bool closing_ = false, aborted_ = false;
  int status_ = BAOREC_OK;
  std::string message_;
  double read_s_ = 0, write_s_ = 0, wait_s_ = 0;
  std::thread reader_, writer_;
};

}  // namespace baorec
