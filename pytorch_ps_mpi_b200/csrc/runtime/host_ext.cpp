// Host-side native runtime of pytorch_ps_mpi_b200 (no CUDA dependency; builds with g++).
//
//  * ShmComm      — single-node message transport over POSIX shared memory: one lock-free
//                   SPSC byte ring per ordered (src,dst) pair, non-blocking post_send /
//                   post_recv (incl. ANY_SOURCE) progressed MPI-style inside wait()/test().
//                   It is the stand-in for libmpi's shared-memory BTL that the reference
//                   reaches through mpi4py (/root/reference/mpi_comms.py:88,132,153,162) and
//                   carries the CPU plumbing configuration (2-rank MLP, BASELINE config 1).
//  * pack_ptrs    — gather raw tensor bytes straight from data_ptr() into one frame, the
//                   finished form of blosc.compress_ptr in /root/reference/serialization.py:22-23.
//  * byteshuffle  — blosc's shuffle filter (element-byte transposition) ahead of deflate.
//
// Messages carry an explicit length header: no sentinel scan, no 10x slots
// (/root/reference/mpi_comms.py:80-85,96-104).
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <atomic>
#include <cerrno>
#include <chrono>
#include <csignal>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <deque>
#include <fcntl.h>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <sys/mman.h>
#include <sys/stat.h>
#include <thread>
#include <unistd.h>
#include <vector>

namespace py = pybind11;

namespace {

constexpr uint32_t kMagic = 0x50534232u;  // 'PSB2'
constexpr uint32_t kMsgMagic = 0x4d534721u;
constexpr int kAnySource = -1;

struct alignas(64) SegHeader {
  std::atomic<uint32_t> magic;
  uint32_t world;
  uint64_t ring_bytes;
  std::atomic<uint32_t> attached;
  std::atomic<uint32_t> barrier_count;
  std::atomic<uint32_t> barrier_gen;
  std::atomic<uint32_t> aborted;
  std::atomic<int32_t> pids[64];
};

struct alignas(64) RingCtl {
  alignas(64) std::atomic<uint64_t> head;  // bytes produced (writer-owned)
  alignas(64) std::atomic<uint64_t> tail;  // bytes consumed (reader-owned)
};

struct MsgHeader {
  uint64_t len;
  uint32_t tag;
  uint32_t magic;
};
static_assert(sizeof(MsgHeader) == 16, "MsgHeader must be 16 bytes");

inline uint64_t pad16(uint64_t n) { return (n + 15) & ~uint64_t(15); }

// A received message; exposes the buffer protocol so Python can build zero-copy views.
struct Message {
  std::unique_ptr<char[]> data;
  uint64_t len = 0;
  int src = -1;
  uint32_t tag = 0;
};

struct Op {
  bool is_send = false;
  int peer = 0;          // dst for sends; src (or kAnySource) for recvs
  uint32_t tag = 0;
  bool done = false;
  // send state
  py::object keepalive;  // owner of the send buffer
  const char* sbuf = nullptr;
  uint64_t slen = 0;
  uint64_t sent = 0;     // bytes of (header+payload+pad) already written
  // recv state
  bool matched = false;
  uint64_t got = 0;
  std::shared_ptr<Message> msg;
};

class ShmComm {
 public:
  ShmComm(const std::string& name, int rank, int world, uint64_t ring_bytes, double attach_timeout_s)
      : name_(name), rank_(rank), world_(world) {
    if (world < 1 || world > 64) throw std::runtime_error("ShmComm: world must be in [1,64]");
    ring_bytes = pad16(ring_bytes < 4096 ? 4096 : ring_bytes);
    const uint64_t nrings = uint64_t(world) * world;
    ctl_off_ = pad16(sizeof(SegHeader));
    data_off_ = ctl_off_ + nrings * sizeof(RingCtl);
    total_ = data_off_ + nrings * ring_bytes;
    int fd = -1;
    if (rank == 0) {
      shm_unlink(name.c_str());
      fd = shm_open(name.c_str(), O_CREAT | O_EXCL | O_RDWR, 0600);
      if (fd < 0) throw std::runtime_error("ShmComm: shm_open(create) failed: " + std::string(strerror(errno)));
      if (ftruncate(fd, (off_t)total_) != 0) {
        close(fd);
        shm_unlink(name.c_str());
        throw std::runtime_error("ShmComm: ftruncate failed: " + std::string(strerror(errno)));
      }
    } else {
      auto t0 = std::chrono::steady_clock::now();
      while (true) {
        fd = shm_open(name.c_str(), O_RDWR, 0600);
        if (fd >= 0) {
          struct stat st;
          if (fstat(fd, &st) == 0 && (uint64_t)st.st_size >= total_) break;
          close(fd);
          fd = -1;
        }
        if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > attach_timeout_s)
          throw std::runtime_error("ShmComm: timed out waiting for rank 0 to create the segment");
        std::this_thread::sleep_for(std::chrono::milliseconds(2));
      }
    }
    base_ = (char*)mmap(nullptr, total_, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (base_ == MAP_FAILED) throw std::runtime_error("ShmComm: mmap failed");
    hdr_ = reinterpret_cast<SegHeader*>(base_);
    if (rank == 0) {
      hdr_->world = world;
      hdr_->ring_bytes = ring_bytes;
      hdr_->attached.store(0);
      hdr_->barrier_count.store(0);
      hdr_->barrier_gen.store(0);
      hdr_->aborted.store(0);
      for (auto& p : hdr_->pids) p.store(0);
      for (uint64_t i = 0; i < nrings; ++i) {
        ctl(i)->head.store(0);
        ctl(i)->tail.store(0);
      }
      hdr_->magic.store(kMagic, std::memory_order_release);
    } else {
      auto t0 = std::chrono::steady_clock::now();
      while (hdr_->magic.load(std::memory_order_acquire) != kMagic) {
        if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > attach_timeout_s)
          throw std::runtime_error("ShmComm: segment never initialised");
        std::this_thread::sleep_for(std::chrono::milliseconds(1));
      }
      if ((int)hdr_->world != world || hdr_->ring_bytes != ring_bytes)
        throw std::runtime_error("ShmComm: world/ring size mismatch with rank 0");
    }
    ring_bytes_ = ring_bytes;
    hdr_->pids[rank].store((int32_t)getpid());
    uint32_t n = hdr_->attached.fetch_add(1) + 1;
    // the last rank to attach removes the name: nothing is left behind if a rank dies later
    if ((int)n == world) shm_unlink(name.c_str());
    sends_.resize(world);
  }

  ~ShmComm() {
    if (base_ && base_ != MAP_FAILED) munmap(base_, total_);
    if (rank_ == 0) shm_unlink(name_.c_str());
  }

  int rank() const { return rank_; }
  int world() const { return world_; }
  uint64_t ring_bytes() const { return ring_bytes_; }

  std::shared_ptr<Op> post_send(int dst, uint32_t tag, py::buffer buf) {
    if (dst < 0 || dst >= world_) throw std::runtime_error("post_send: bad destination");
    py::buffer_info info = buf.request();
    auto op = std::make_shared<Op>();
    op->is_send = true;
    op->peer = dst;
    op->tag = tag;
    op->keepalive = buf;
    op->sbuf = static_cast<const char*>(info.ptr);
    op->slen = (uint64_t)info.size * (uint64_t)info.itemsize;
    {
      std::lock_guard<std::mutex> g(mu_);
      sends_[dst].push_back(op);
    }
    return op;
  }

  std::shared_ptr<Op> post_recv(int src, uint32_t tag) {
    if (src != kAnySource && (src < 0 || src >= world_)) throw std::runtime_error("post_recv: bad source");
    auto op = std::make_shared<Op>();
    op->is_send = false;
    op->peer = src;
    op->tag = tag;
    {
      std::lock_guard<std::mutex> g(mu_);
      recvs_.push_back(op);
    }
    return op;
  }

  // One pass over all pending operations.  Returns true if any byte moved.
  bool progress() {
    std::lock_guard<std::mutex> g(mu_);
    bool moved = false;
    // ---- sends: per destination FIFO (the ring is a byte stream) ----
    for (int d = 0; d < world_; ++d) {
      auto& q = sends_[d];
      while (!q.empty()) {
        Op& op = *q.front();
        RingCtl* c = ctl(idx(rank_, d));
        char* data = ring(idx(rank_, d));
        const uint64_t total = sizeof(MsgHeader) + pad16(op.slen);
        uint64_t head = c->head.load(std::memory_order_relaxed);
        uint64_t tail = c->tail.load(std::memory_order_acquire);
        uint64_t space = ring_bytes_ - (head - tail);
        if (space == 0) break;
        uint64_t n = std::min(space, total - op.sent);
        // source bytes come from three regions: header, payload, zero pad
        uint64_t w = 0;
        MsgHeader h{op.slen, op.tag, kMsgMagic};
        while (w < n) {
          uint64_t pos = op.sent + w;
          const char* src;
          uint64_t avail;
          static const char zeros[16] = {0};
          if (pos < sizeof(MsgHeader)) {
            src = reinterpret_cast<const char*>(&h) + pos;
            avail = sizeof(MsgHeader) - pos;
          } else if (pos < sizeof(MsgHeader) + op.slen) {
            src = op.sbuf + (pos - sizeof(MsgHeader));
            avail = sizeof(MsgHeader) + op.slen - pos;
          } else {
            src = zeros;
            avail = total - pos;
          }
          uint64_t chunk = std::min(avail, n - w);
          copy_in(data, (head + w) % ring_bytes_, src, chunk);
          w += chunk;
        }
        op.sent += n;
        c->head.store(head + n, std::memory_order_release);
        moved = moved || n > 0;
        if (op.sent == total) {
          op.done = true;
          op.keepalive = py::object();  // needs the GIL: progress() is only called with it held... see wait()
          q.pop_front();
        } else {
          break;
        }
      }
    }
    // ---- receives: per source stream parsing ----
    for (int s = 0; s < world_; ++s) {
      RingCtl* c = ctl(idx(s, rank_));
      char* data = ring(idx(s, rank_));
      while (true) {
        uint64_t tail = c->tail.load(std::memory_order_relaxed);
        uint64_t head = c->head.load(std::memory_order_acquire);
        uint64_t avail = head - tail;
        std::shared_ptr<Op>& cur = inflight_[s];
        if (!cur) {
          if (avail < sizeof(MsgHeader)) break;
          MsgHeader h;
          copy_out(reinterpret_cast<char*>(&h), data, tail % ring_bytes_, sizeof(MsgHeader));
          if (h.magic != kMsgMagic) throw std::runtime_error("ShmComm: corrupted message header");
          // match: first posted recv with (src == s or ANY) and equal tag, not yet matched
          std::shared_ptr<Op> hit;
          for (auto& r : recvs_) {
            if (!r->matched && (r->peer == s || r->peer == kAnySource) && r->tag == h.tag) {
              hit = r;
              break;
            }
          }
          if (!hit) break;  // nobody wants it yet: leave it in the ring
          hit->matched = true;
          hit->msg = std::make_shared<Message>();
          hit->msg->len = h.len;
          hit->msg->src = s;
          hit->msg->tag = h.tag;
          hit->msg->data.reset(new char[h.len ? h.len : 1]);
          hit->got = 0;
          cur = hit;
          c->tail.store(tail + sizeof(MsgHeader), std::memory_order_release);
          moved = true;
          continue;
        }
        Op& op = *cur;
        const uint64_t padded = pad16(op.msg->len);
        if (op.got < padded) {
          if (avail == 0) break;
          uint64_t n = std::min(avail, padded - op.got);
          uint64_t useful = op.got < op.msg->len ? std::min(n, op.msg->len - op.got) : 0;
          if (useful) copy_out(op.msg->data.get() + op.got, data, tail % ring_bytes_, useful);
          op.got += n;
          c->tail.store(tail + n, std::memory_order_release);
          moved = true;
        }
        if (op.got == padded) {
          op.done = true;
          for (auto it = recvs_.begin(); it != recvs_.end(); ++it)
            if (it->get() == &op) {
              recvs_.erase(it);
              break;
            }
          cur.reset();
        } else {
          break;
        }
      }
    }
    return moved;
  }

  // Block until `op` completes.  The GIL is held while progress() runs (it may drop Python
  // references) and released while we back off, so peers in the same interpreter (threads) and
  // signal handlers keep running.
  void wait(const std::shared_ptr<Op>& op, double timeout_s) {
    auto t0 = std::chrono::steady_clock::now();
    uint64_t idle = 0;
    while (!op->done) {
      bool moved = progress();
      if (op->done) break;
      if (moved) {
        idle = 0;
        continue;
      }
      ++idle;
      if (idle > 64) {
        py::gil_scoped_release rel;
        if (idle > 2000)
          std::this_thread::sleep_for(std::chrono::microseconds(50));
        else
          std::this_thread::yield();
      }
      if ((idle & 1023) == 0) {
        if (hdr_->aborted.load()) throw std::runtime_error("ShmComm: job aborted by a peer");
        check_peers_alive();
        if (timeout_s > 0 &&
            std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > timeout_s)
          throw std::runtime_error("ShmComm: wait timed out (peer stalled or dead)");
        if (PyErr_CheckSignals() != 0) throw py::error_already_set();
      }
    }
  }

  bool test(const std::shared_ptr<Op>& op) {
    if (!op->done) progress();
    return op->done;
  }

  void barrier(double timeout_s) {
    if (world_ == 1) return;
    auto t0 = std::chrono::steady_clock::now();
    uint32_t gen = hdr_->barrier_gen.load(std::memory_order_acquire);
    if (hdr_->barrier_count.fetch_add(1, std::memory_order_acq_rel) + 1 == (uint32_t)world_) {
      hdr_->barrier_count.store(0, std::memory_order_relaxed);
      hdr_->barrier_gen.store(gen + 1, std::memory_order_release);
      return;
    }
    uint64_t idle = 0;
    while (hdr_->barrier_gen.load(std::memory_order_acquire) == gen) {
      progress();  // keep draining so peers blocked on a full ring can reach the barrier
      if (++idle > 64) {
        py::gil_scoped_release rel;
        std::this_thread::sleep_for(std::chrono::microseconds(20));
      }
      if ((idle & 1023) == 0) {
        if (hdr_->aborted.load()) throw std::runtime_error("ShmComm: job aborted by a peer");
        check_peers_alive();
        if (timeout_s > 0 &&
            std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > timeout_s)
          throw std::runtime_error("ShmComm: barrier timed out");
      }
    }
  }

  void abort() { hdr_->aborted.store(1); }

  // A peer is gone if its pid no longer exists OR it is a zombie (exited, not yet reaped by its parent —
  // kill(pid, 0) still succeeds for those).
  static bool pid_alive(int32_t pid) {
    if (kill(pid, 0) != 0 && errno == ESRCH) return false;
    char path[64], buf[256];
    snprintf(path, sizeof(path), "/proc/%d/stat", (int)pid);
    FILE* f = fopen(path, "r");
    if (!f) return errno != ENOENT;   // no /proc (or no permission): trust kill()
    bool alive = true;
    if (fgets(buf, sizeof(buf), f)) {
      const char* rp = strrchr(buf, ')');           // "pid (comm) S ..."
      if (rp && rp[1] == ' ' && (rp[2] == 'Z' || rp[2] == 'X')) alive = false;
    }
    fclose(f);
    return alive;
  }

  std::vector<int> dead_peers() const {
    std::vector<int> dead;
    for (int r = 0; r < world_; ++r) {
      int32_t pid = hdr_->pids[r].load();
      if (r != rank_ && pid > 0 && !pid_alive(pid)) dead.push_back(r);
    }
    return dead;
  }

 private:
  void check_peers_alive() const {
    auto dead = dead_peers();
    if (!dead.empty())
      throw std::runtime_error("ShmComm: peer rank " + std::to_string(dead[0]) + " is no longer running");
  }
  uint64_t idx(int s, int d) const { return uint64_t(s) * world_ + d; }
  RingCtl* ctl(uint64_t i) const { return reinterpret_cast<RingCtl*>(base_ + ctl_off_) + i; }
  char* ring(uint64_t i) const { return base_ + data_off_ + i * ring_bytes_; }
  void copy_in(char* ringp, uint64_t pos, const char* src, uint64_t n) const {
    uint64_t first = std::min(n, ring_bytes_ - pos);
    memcpy(ringp + pos, src, first);
    if (n > first) memcpy(ringp, src + first, n - first);
  }
  void copy_out(char* dst, const char* ringp, uint64_t pos, uint64_t n) const {
    uint64_t first = std::min(n, ring_bytes_ - pos);
    memcpy(dst, ringp + pos, first);
    if (n > first) memcpy(dst + first, ringp, n - first);
  }

  std::string name_;
  int rank_, world_;
  uint64_t ring_bytes_ = 0, ctl_off_ = 0, data_off_ = 0, total_ = 0;
  char* base_ = nullptr;
  SegHeader* hdr_ = nullptr;
  std::mutex mu_;
  std::vector<std::deque<std::shared_ptr<Op>>> sends_;
  std::deque<std::shared_ptr<Op>> recvs_;
  std::shared_ptr<Op> inflight_[64];
};

// ---- tensor packing & blosc-style byte shuffle ----------------------------------------

void pack_ptrs(py::buffer dst, const std::vector<uint64_t>& offs, const std::vector<uint64_t>& ptrs,
               const std::vector<uint64_t>& sizes) {
  py::buffer_info info = dst.request(true);
  char* base = static_cast<char*>(info.ptr);
  const uint64_t cap = (uint64_t)info.size * (uint64_t)info.itemsize;
  if (offs.size() != ptrs.size() || offs.size() != sizes.size()) throw std::runtime_error("pack_ptrs: length mismatch");
  for (size_t i = 0; i < offs.size(); ++i)
    if (offs[i] + sizes[i] > cap) throw std::runtime_error("pack_ptrs: destination overflow");
  py::gil_scoped_release rel;
  uint64_t total = 0;
  for (auto s : sizes) total += s;
  const unsigned hw = std::max(1u, std::min(8u, std::thread::hardware_concurrency()));
  if (total < (8u << 20) || hw == 1) {
    for (size_t i = 0; i < offs.size(); ++i)
      if (sizes[i]) memcpy(base + offs[i], reinterpret_cast<const void*>(ptrs[i]), sizes[i]);
    return;
  }
  // large frames: split every tensor into 1 MiB chunks and fan out over threads
  struct Chunk { char* d; const char* s; uint64_t n; };
  std::vector<Chunk> chunks;
  for (size_t i = 0; i < offs.size(); ++i)
    for (uint64_t o = 0; o < sizes[i]; o += (1u << 20))
      chunks.push_back({base + offs[i] + o, reinterpret_cast<const char*>(ptrs[i]) + o,
                        std::min<uint64_t>(1u << 20, sizes[i] - o)});
  std::atomic<size_t> next{0};
  std::vector<std::thread> th;
  for (unsigned t = 0; t < hw; ++t)
    th.emplace_back([&] {
      for (size_t c; (c = next.fetch_add(1)) < chunks.size();) memcpy(chunks[c].d, chunks[c].s, chunks[c].n);
    });
  for (auto& t : th) t.join();
}

py::bytes byteshuffle(py::buffer src, int typesize) {
  py::buffer_info info = src.request();
  const uint8_t* in = static_cast<const uint8_t*>(info.ptr);
  const uint64_t n = (uint64_t)info.size * (uint64_t)info.itemsize;
  if (typesize < 1) throw std::runtime_error("byteshuffle: typesize must be >= 1");
  std::string out(n, '\0');
  const uint64_t ne = n / typesize, body = ne * typesize;
  uint8_t* o = reinterpret_cast<uint8_t*>(&out[0]);
  {
    py::gil_scoped_release rel;
    for (int b = 0; b < typesize; ++b) {
      uint8_t* dst = o + uint64_t(b) * ne;
      const uint8_t* s = in + b;
      for (uint64_t e = 0; e < ne; ++e) dst[e] = s[e * typesize];
    }
    memcpy(o + body, in + body, n - body);
  }
  return py::bytes(out);
}

py::bytearray byteunshuffle(py::buffer src, int typesize) {
  py::buffer_info info = src.request();
  const uint8_t* in = static_cast<const uint8_t*>(info.ptr);
  const uint64_t n = (uint64_t)info.size * (uint64_t)info.itemsize;
  if (typesize < 1) throw std::runtime_error("byteunshuffle: typesize must be >= 1");
  std::string out(n, '\0');
  const uint64_t ne = n / typesize, body = ne * typesize;
  uint8_t* o = reinterpret_cast<uint8_t*>(&out[0]);
  {
    py::gil_scoped_release rel;
    for (int b = 0; b < typesize; ++b) {
      const uint8_t* s = in + uint64_t(b) * ne;
      uint8_t* dst = o + b;
      for (uint64_t e = 0; e < ne; ++e) dst[e * typesize] = s[e];
    }
    memcpy(o + body, in + body, n - body);
  }
  return py::bytearray(out);
}

}  // namespace

PYBIND11_MODULE(_psb200_host, m) {
  m.doc() = "pytorch_ps_mpi_b200 host runtime: shm transport, tensor packing, byte shuffle";
  m.attr("ANY_SOURCE") = kAnySource;

  py::class_<Message, std::shared_ptr<Message>>(m, "Message", py::buffer_protocol())
      .def_readonly("src", &Message::src)
      .def_readonly("tag", &Message::tag)
      .def("__len__", [](const Message& s) { return s.len; })
      .def_buffer([](Message& s) {
        return py::buffer_info(s.data.get(), 1, py::format_descriptor<uint8_t>::format(), 1,
                               {(py::ssize_t)s.len}, {(py::ssize_t)1}, /*readonly=*/false);
      });

  py::class_<Op, std::shared_ptr<Op>>(m, "Op")
      .def_readonly("done", &Op::done)
      .def_readonly("is_send", &Op::is_send)
      .def_property_readonly("message", [](const Op& o) { return o.msg; });

  py::class_<ShmComm>(m, "ShmComm")
      .def(py::init<const std::string&, int, int, uint64_t, double>(), py::arg("name"), py::arg("rank"),
           py::arg("world"), py::arg("ring_bytes") = (4u << 20), py::arg("attach_timeout_s") = 60.0)
      .def_property_readonly("rank", &ShmComm::rank)
      .def_property_readonly("world", &ShmComm::world)
      .def_property_readonly("ring_bytes", &ShmComm::ring_bytes)
      .def("post_send", &ShmComm::post_send, py::arg("dst"), py::arg("tag"), py::arg("buf"))
      .def("post_recv", &ShmComm::post_recv, py::arg("src"), py::arg("tag"))
      .def("progress", &ShmComm::progress)
      .def("wait", &ShmComm::wait, py::arg("op"), py::arg("timeout_s") = 0.0)
      .def("test", &ShmComm::test)
      .def("barrier", &ShmComm::barrier, py::arg("timeout_s") = 0.0)
      .def("abort", &ShmComm::abort)
      .def("dead_peers", &ShmComm::dead_peers);

  m.def("pack_ptrs", &pack_ptrs, "memcpy raw tensor bytes (by pointer) into a frame");
  m.def("byteshuffle", &byteshuffle);
  m.def("byteunshuffle", &byteunshuffle);
}
