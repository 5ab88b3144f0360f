// File-to-file front end ("feature .bin reader/writer + streaming front-end").
//
// The data path of the reference's command-line driver, src/cpp/dnn.cc:55-78 —
//   BatchData input(path)  →  CalculationContext::Calculate(input)  →  output->dumpToFile(path, binary)
// — holds the whole feature matrix and the whole score matrix in memory (float_dnn.cc:85-105 reads every
// float of the file into one array, dnn.cc:451 allocates n × outputs floats).  For BASELINE config 4
// (one million frames: 1.76 GB of features, 32 GB of scores) that is not an option, so here the
// two files are streamed: a reader thread fills chunks from the big-endian feature file, the CALLING
// thread pushes each chunk through fdnn_calculate_sink (host → device, kernels, device → host through the
// model's pooled page-locked staging) and appends the scores to the dump piece by piece as they land.
// Only the public C ABI is used below (include/fdnn.h); all arithmetic is the library's GPU path — a
// model handle cannot even exist without a usable GPU, and there is no CPU path.
//
// File formats (both follow the reference exactly):
//   features   big-endian int32 frames, int32 dim, fp32 rows       float_dnn.cc:85-105, BatchData.java:80-91
//   BIN dump   native-endian uint32 frames, uint32 dim, fp32 rows   float_dnn.cc:128-164 (binary = true)
//   TXT dump   one row per line, values printed by `ostream << float` (= "%g", six significant digits)
//              and separated by one blank                           float_dnn.cc:128-164 (binary = false)

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <new>
#include <string>
#include <system_error>
#include <sys/types.h>
#include <thread>
#include <vector>

#include "../../include/fdnn.h"
#include "fdnn_internal.h"

namespace {

using fdnn::set_error;

// slot indices handed from one pipeline stage to the next
class Channel {
 public:
  void push(int v) {
    {
      std::lock_guard<std::mutex> lk(mu_);
      q_.push_back(v);
    }
    cv_.notify_one();
  }
  // false: the channel was closed and is empty
  bool pop(int *v) {
    std::unique_lock<std::mutex> lk(mu_);
    cv_.wait(lk, [&] { return !q_.empty() || closed_; });
    if (q_.empty()) return false;
    *v = q_.front();
    q_.pop_front();
    return true;
  }
  void close() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      closed_ = true;
    }
    cv_.notify_all();
  }

 private:
  std::mutex mu_;
  std::condition_variable cv_;
  std::deque<int> q_;
  bool closed_ = false;
};

// the first failure of any stage; its text reaches fdnn_last_error() of the calling thread at the end
struct Failure {
  std::mutex mu;
  int rc = FDNN_OK;
  std::string text;
  std::atomic<bool> any{false};
  void set(int code, const std::string &msg) {
    std::lock_guard<std::mutex> lk(mu);
    if (rc == FDNN_OK) {
      rc = code;
      text = msg;
    }
    any.store(true, std::memory_order_release);
  }
};

uint32_t be32(const uint8_t *p) { return (uint32_t(p[0]) << 24) | (uint32_t(p[1]) << 16) | (uint32_t(p[2]) << 8) | uint32_t(p[3]); }

// `ostream << float` of the reference's text dump: printf's %g with the stream's default precision of 6
void append_row_txt(std::string &line, const float *row, int dim) {
  char buf[32];
  for (int j = 0; j < dim; ++j) {
    const int k = std::snprintf(buf, sizeof buf, "%g", double(row[j]));
    line.append(buf, size_t(k));
    if (j + 1 < dim) line.push_back(' ');
  }
  line.push_back('\n');
}

bool write_rows(std::FILE *f, int format, const float *rows, int frames, int dim, std::string &scratch) {
  if (format == FDNN_DUMP_BIN) return std::fwrite(rows, sizeof(float) * size_t(dim), size_t(frames), f) == size_t(frames);
  for (int r = 0; r < frames; ++r) {
    scratch.clear();
    append_row_txt(scratch, rows + size_t(r) * size_t(dim), dim);
    if (std::fwrite(scratch.data(), 1, scratch.size(), f) != scratch.size()) return false;
  }
  return true;
}

bool write_header(std::FILE *f, int format, long long frames, int dim) {
  if (format != FDNN_DUMP_BIN) return true;  // the text dump has no header
  const uint32_t hdr[2] = {uint32_t(frames), uint32_t(dim)};
  return std::fwrite(hdr, sizeof hdr, 1, f) == 1;
}

double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

struct FileCloser {
  std::FILE *f = nullptr;
  ~FileCloser() {
    if (f) std::fclose(f);
  }
};

}  // namespace

extern "C" {

int fdnn_output_dump_write_txt(const char *path, const float *data, int frames, int dim) {
  if (!path || (!data && frames > 0) || frames < 0 || dim <= 0) {
    set_error("bad argument to fdnn_output_dump_write_txt");
    return FDNN_EINVAL;
  }
  FileCloser out;
  out.f = std::fopen(path, "wb");
  if (!out.f) {
    set_error(std::string("cannot open for writing: ") + path);
    return FDNN_EIO;
  }
  std::string scratch;
  const bool ok = write_rows(out.f, FDNN_DUMP_TXT, data, frames, dim, scratch);
  const bool closed = std::fclose(out.f) == 0;
  out.f = nullptr;
  if (!ok || !closed) {
    set_error(std::string("short write: ") + path);
    return FDNN_EIO;
  }
  return FDNN_OK;
}

int fdnn_calculate_file(fdnn_model *model, const char *feature_bin_path, const char *out_path, int out_format, int chunk_frames,
                        long long *frames_done) {
  if (frames_done) *frames_done = 0;
  if (!model || !feature_bin_path || !out_path || (out_format != FDNN_DUMP_BIN && out_format != FDNN_DUMP_TXT) || chunk_frames < 0) {
    set_error("bad argument to fdnn_calculate_file");
    return FDNN_EINVAL;
  }
  const int I = fdnn_input_dim(model), O = fdnn_output_dim(model);
  if (I <= 0 || O <= 0) {
    set_error("bad model handle");
    return FDNN_EINVAL;
  }

  // ---- headers, on the calling thread: a bad file fails before anything is written ----
  FileCloser in, out;
  in.f = std::fopen(feature_bin_path, "rb");
  if (!in.f) {
    set_error(std::string("cannot open feature file: ") + feature_bin_path);
    return FDNN_EIO;
  }
  uint8_t hdr[8];
  if (std::fread(hdr, 1, 8, in.f) != 8) {
    set_error(std::string("feature file shorter than its header: ") + feature_bin_path);
    return FDNN_EIO;
  }
  const long long n = (long long) int32_t(be32(hdr));
  const int d = int(int32_t(be32(hdr + 4)));
  if (n < 0 || d <= 0) {
    set_error(std::string("not a usable feature file: ") + feature_bin_path);
    return FDNN_EFORMAT;
  }
  // The network's input width is the file's padded to a multiple of four with zero weights (float_dnn.cc:32-33, 61-66);
  // features may come either way: already padded, or as wide as the unpadded network (missing columns are zeros).
  if (d > I || I - d >= 4) {
    set_error("feature dimension " + std::to_string(d) + " does not match the network's input dimension " + std::to_string(I));
    return FDNN_EINVAL;
  }
  if (std::fseek(in.f, 0, SEEK_END) == 0) {  // (not seekable, e.g. a pipe: a short file shows up as a short read below)
    const long long size = (long long) std::ftell(in.f);
    if (size >= 0 && size - 8 < n * (long long) d * 4) {
      set_error(std::string("feature file is shorter than its header announces: ") + feature_bin_path);
      return FDNN_EIO;
    }
    std::fseek(in.f, 8, SEEK_SET);
  }
  const int n_dev = fdnn_device_count(model);
  const int chunk = chunk_frames > 0 ? chunk_frames : (out_format == FDNN_DUMP_BIN ? 4096 : 1024) * (n_dev > 0 ? n_dev : 1);
  out.f = std::fopen(out_path, "wb");
  if (!out.f) {
    set_error(std::string("cannot open for writing: ") + out_path);
    return FDNN_EIO;
  }
  if (!write_header(out.f, out_format, n, O)) {
    set_error(std::string("short write: ") + out_path);
    return FDNN_EIO;
  }

  // FDNN_FILE_DEBUG=1: where the time of a call goes (stderr), per stage the time spent working, not waiting
  const char *dbg_env = std::getenv("FDNN_FILE_DEBUG");
  const bool dbg = dbg_env && dbg_env[0] == '1';
  const double t_begin = now_ms();
  double ms_read = 0.0, ms_swap = 0.0, ms_gpu = 0.0, ms_write = 0.0;

  // ---- reader thread → this thread (GPU + dump) ----
  // Two plain chunk buffers: while one is on its way through the GPU the reader fills the other.  The library's own pooled
  // page-locked staging does the rest (fdnn_calculate_sink: upload, kernels and download of 512-frame passes overlap, and the
  // scores are written to the dump straight out of the transfer buffer, 128 frames at a time, while later ones are still
  // crossing PCIe) — buffers of our own would have to be page-locked per call, which costs more than a short file takes
  // (measured on the B200 host: 120-460 ms to lock 207 MB, as much again to release them).
  constexpr int kSlots = 2;
  struct Slot {
    std::vector<float> in;  // [chunk][I]
    int frames = 0;
  } slot[kSlots];
  const int n_slots = int(std::min<long long>(kSlots, std::max<long long>(1, (n + chunk - 1) / chunk)));
  const size_t cap_frames = size_t(std::min<long long>(chunk, std::max<long long>(n, 1)));
  std::vector<uint8_t> raw;       // the reader's view of the file: big-endian rows of d floats
  std::vector<float> text_rows;   // text dump: the scores of the chunk on the GPU
  try {  // everything is allocated before the reader thread exists; nothing may throw across the C ABI
    for (int i = 0; i < n_slots; ++i) slot[i].in.resize(cap_frames * size_t(I));
    raw.resize(cap_frames * size_t(d) * 4);
    if (out_format == FDNN_DUMP_TXT) text_rows.resize(cap_frames * size_t(O));
  } catch (const std::bad_alloc &) {
    set_error("out of host memory for chunks of " + std::to_string(cap_frames) + " frames");
    return FDNN_ENOMEM;
  }
  Channel to_reader, to_compute;
  Failure failure;
  for (int i = 0; i < n_slots; ++i) to_reader.push(i);

  auto read_chunks = [&] {
    long long done = 0;
    int s = 0;
    while (done < n && !failure.any.load(std::memory_order_acquire) && to_reader.pop(&s)) {
      const int frames = int(std::min<long long>(chunk, n - done));
      const size_t words = size_t(frames) * size_t(d);
      const double t0 = now_ms();
      if (std::fread(raw.data(), 4, words, in.f) != words) {
        failure.set(FDNN_EIO, std::string("short read: ") + feature_bin_path);
        break;
      }
      const double t1 = now_ms();
      float *dst = slot[s].in.data();
      const uint8_t *src = raw.data();
      for (int r = 0; r < frames; ++r) {
        uint32_t *row = reinterpret_cast<uint32_t *>(dst + size_t(r) * size_t(I));
        for (int k = 0; k < d; ++k, src += 4) row[k] = be32(src);
        for (int k = d; k < I; ++k) row[k] = 0u;
      }
      ms_read += t1 - t0;
      ms_swap += now_ms() - t1;
      slot[s].frames = frames;
      done += frames;
      to_compute.push(s);
    }
    to_compute.close();
  };
  std::thread reader;
  try {
    reader = std::thread(read_chunks);
  } catch (const std::system_error &) {
    set_error("cannot start the reader thread");
    return FDNN_ENOMEM;
  }

  // The sink sees the pieces of ONE device in frame order; the devices of a group deliver their shards of a call interleaved
  // (fdnn_api.cu: calculate_impl collects round-robin).  Binary dump: a piece that is not the continuation of the previous one
  // is written at its own offset.  Text dump (rows have no fixed length): the rows of a chunk are collected and formatted in
  // order once the call has returned.
  struct DumpSink {
    std::FILE *f;
    int format, dim;
    long long base = 0;      // frames of the chunks before this one
    long long cursor = 0;    // frame the file position stands at (binary)
    long long written = 0;
    double ms = 0.0;
    bool io_error = false;
    float *rows;  // text: this chunk's scores
    std::string scratch;
  } sink{out.f, out_format, O, 0, 0, 0, 0.0, false, text_rows.data(), std::string()};
  const fdnn_sink_fn take_piece = [](void *user, int first_frame, int n_frames, const float *rows) -> int {
    auto *ds = static_cast<DumpSink *>(user);  // called on this thread
    const double t0 = now_ms();
    bool ok = true;
    if (ds->format == FDNN_DUMP_BIN) {
      const long long at = ds->base + first_frame;
      if (at != ds->cursor) ok = fseeko(ds->f, off_t(8 + at * (long long) ds->dim * 4), SEEK_SET) == 0;
      ok = ok && write_rows(ds->f, FDNN_DUMP_BIN, rows, n_frames, ds->dim, ds->scratch);
      ds->cursor = at + n_frames;
      if (ok) ds->written += n_frames;
    } else {
      std::memcpy(ds->rows + size_t(first_frame) * size_t(ds->dim), rows, size_t(n_frames) * size_t(ds->dim) * sizeof(float));
    }
    ds->ms += now_ms() - t0;
    if (!ok) ds->io_error = true;
    return ok ? 0 : 1;
  };
  {
    int s = 0;
    while (to_compute.pop(&s)) {
      if (!failure.any.load(std::memory_order_acquire)) {
        const double t0 = now_ms();
        const int rc = fdnn_calculate_sink(model, slot[s].in.data(), slot[s].frames, I, take_piece, &sink);
        if (rc == FDNN_OK && out_format == FDNN_DUMP_TXT) {
          const double t1 = now_ms();
          if (write_rows(out.f, FDNN_DUMP_TXT, sink.rows, slot[s].frames, O, sink.scratch))
            sink.written += slot[s].frames;
          else
            sink.io_error = true;
          sink.ms += now_ms() - t1;
        }
        sink.base += slot[s].frames;
        ms_gpu += now_ms() - t0;
        if (sink.io_error)
          failure.set(FDNN_EIO, std::string("short write: ") + out_path);
        else if (rc != FDNN_OK)
          failure.set(rc, fdnn_last_error());
      }
      if (failure.any.load(std::memory_order_acquire)) {
        to_reader.close();  // the reader may be waiting for a free buffer
        continue;           // keep draining so that the reader's pushes are consumed and it can finish
      }
      to_reader.push(s);
    }
  }
  reader.join();
  to_reader.close();
  ms_write = sink.ms;

  const bool closed = std::fclose(out.f) == 0;
  out.f = nullptr;
  if (frames_done) *frames_done = sink.written;
  if (dbg)
    std::fprintf(stderr, "fdnn_calculate_file: %lld frames, chunks of %d: %.1f ms (reader: fread %.1f + byte swap %.1f ms busy; this thread: "
                         "fdnn_calculate_sink %.1f ms, of which writing the dump %.1f)\n",
                 n, chunk, now_ms() - t_begin, ms_read, ms_swap, ms_gpu, ms_write);
  if (failure.rc != FDNN_OK) {
    set_error(failure.text);
    return failure.rc;
  }
  if (!closed) {
    set_error(std::string("short write: ") + out_path);
    return FDNN_EIO;
  }
  return FDNN_OK;
}

}  // extern "C"
