// TEST INFRASTRUCTURE ONLY (oracle build): a tiny MPI stand-in so that the
// UNMODIFIED reference CPU path (/root/reference, MPI + libnuma) can be built
// and run in an image that has neither.  It implements exactly the MPI surface
// the reference touches (grep over core/ comm/ toolkits/ dep/: 12 functions,
// 8 datatypes, 3 reduction ops) and nothing more.
//
//   * world size 1 (default): everything stays inside the process.  The
//     reference still sends to itself (core/graph.hpp:1328-1412 shuffles edges
//     to their owner while a receiver thread probes), hence the mailbox.
//   * world size P>1: launch P processes with NTS_SHIM_SIZE=P, NTS_SHIM_RANK=r
//     and a shared scratch directory NTS_SHIM_DIR.  A message to another rank
//     is a file renamed into `<dir>/inbox_<dst>/`; per-sender sequence numbers
//     keep MPI's non-overtaking order.  Collectives are built on send/recv.
//     This is slow and only meant for Cora-sized oracle runs that pin the
//     reference's partition / chunk / mirror artefacts at P = 2, 4, 8.
#pragma once
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <dirent.h>
#include <mutex>
#include <string>
#include <sys/stat.h>
#include <unistd.h>
#include <vector>

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
struct MPI_Status {
  int MPI_SOURCE;
  int MPI_TAG;
  int MPI_ERROR;
  long shim_bytes;
};

#define MPI_COMM_WORLD 0
#define MPI_STATUS_IGNORE ((MPI_Status *)0)
#define MPI_IN_PLACE ((void *)1)
#define MPI_ANY_SOURCE (-1)
#define MPI_SUCCESS 0

enum {
  MPI_CHAR = 1,
  MPI_UNSIGNED_CHAR,
  MPI_INT,
  MPI_UNSIGNED,
  MPI_LONG,
  MPI_UNSIGNED_LONG,
  MPI_FLOAT,
  MPI_DOUBLE
};
enum { MPI_SUM = 1, MPI_MAX, MPI_MIN };
enum {
  MPI_THREAD_SINGLE,
  MPI_THREAD_FUNNELED,
  MPI_THREAD_SERIALIZED,
  MPI_THREAD_MULTIPLE
};

namespace nts_mpi_shim {

static const int kCollectiveTag = 0x7fff0001;

inline size_t type_size(MPI_Datatype t) {
  switch (t) {
  case MPI_CHAR:
  case MPI_UNSIGNED_CHAR:
    return 1;
  case MPI_INT:
  case MPI_UNSIGNED:
  case MPI_FLOAT:
    return 4;
  default:
    return 8;
  }
}

struct LocalMsg {
  int tag;
  std::vector<char> payload;
};

struct World {
  int rank = 0;
  int size = 1;
  std::string dir;
  double timeout_s = 900.0;
  std::atomic<unsigned long long> next_seq{0};
  // messages to self
  std::mutex mu;
  std::condition_variable cv;
  std::deque<LocalMsg> self_box;
  World() {
    if (const char *e = getenv("NTS_SHIM_SIZE"))
      size = std::max(1, atoi(e));
    if (const char *e = getenv("NTS_SHIM_RANK"))
      rank = atoi(e);
    if (const char *e = getenv("NTS_SHIM_TIMEOUT"))
      timeout_s = atof(e);
    if (size > 1) {
      const char *d = getenv("NTS_SHIM_DIR");
      if (!d) {
        fprintf(stderr, "mpi shim: NTS_SHIM_DIR must be set when NTS_SHIM_SIZE>1\n");
        abort();
      }
      dir = d;
      for (int r = 0; r < size; r++) {
        std::string box = dir + "/inbox_" + std::to_string(r);
        mkdir(box.c_str(), 0777); // EEXIST is fine: every rank tries
      }
    }
  }
};

inline World &world() {
  static World w;
  return w;
}

inline double now_s() {
  return std::chrono::duration<double>(
             std::chrono::steady_clock::now().time_since_epoch())
      .count();
}

// ---- remote (file) mailbox -------------------------------------------------
struct RemoteHit {
  bool found = false;
  int src = -1;
  unsigned long long seq = 0;
  std::string path;
  long bytes = 0;
};

// Oldest (lowest sequence number) pending file from `source` (or any source)
// carrying `tag`.
inline RemoteHit scan_inbox(int source, int tag) {
  World &w = world();
  RemoteHit best;
  std::string box = w.dir + "/inbox_" + std::to_string(w.rank);
  DIR *d = opendir(box.c_str());
  if (!d)
    return best;
  while (struct dirent *ent = readdir(d)) {
    int src, t;
    unsigned long long seq;
    if (sscanf(ent->d_name, "m_%d_%d_%llu", &src, &t, &seq) != 3)
      continue;
    if (t != tag)
      continue;
    if (source != MPI_ANY_SOURCE && src != source)
      continue;
    bool better;
    if (!best.found) {
      better = true;
    } else if (src == best.src) {
      better = seq < best.seq;
    } else {
      better = false; // keep the first source seen; order across sources is free
    }
    if (better) {
      best.found = true;
      best.src = src;
      best.seq = seq;
      best.path = box + "/" + ent->d_name;
    }
  }
  closedir(d);
  if (best.found) {
    struct stat st;
    if (stat(best.path.c_str(), &st) != 0) {
      best.found = false;
    } else {
      best.bytes = (long)st.st_size;
    }
  }
  return best;
}

inline void send_remote(const void *buf, size_t bytes, int dst, int tag) {
  World &w = world();
  unsigned long long seq = w.next_seq.fetch_add(1);
  std::string box = w.dir + "/inbox_" + std::to_string(dst);
  char name[128];
  snprintf(name, sizeof(name), "m_%d_%d_%020llu", w.rank, tag, seq);
  std::string tmp = box + "/.tmp_" + std::to_string(w.rank) + "_" + std::to_string(seq);
  FILE *f = fopen(tmp.c_str(), "wb");
  if (!f) {
    fprintf(stderr, "mpi shim: cannot write %s\n", tmp.c_str());
    abort();
  }
  if (bytes)
    fwrite(buf, 1, bytes, f);
  fclose(f);
  std::string fin = box + "/" + name;
  if (rename(tmp.c_str(), fin.c_str()) != 0) {
    fprintf(stderr, "mpi shim: rename failed for %s\n", fin.c_str());
    abort();
  }
}

// ---- matching ----------------------------------------------------------------
// Blocks until a message matching (source, tag) exists.  When `consume` is set
// the payload is copied into buf (at most cap bytes) and the message removed.
inline void match(int source, int tag, bool consume, void *buf, size_t cap,
                  MPI_Status *st) {
  World &w = world();
  double t0 = now_s();
  for (;;) {
    if (source == MPI_ANY_SOURCE || source == w.rank) {
      std::unique_lock<std::mutex> lk(w.mu);
      for (auto it = w.self_box.begin(); it != w.self_box.end(); ++it) {
        if (it->tag != tag)
          continue;
        if (st) {
          st->MPI_SOURCE = w.rank;
          st->MPI_TAG = tag;
          st->MPI_ERROR = 0;
          st->shim_bytes = (long)it->payload.size();
        }
        if (consume) {
          size_t n = std::min(cap, it->payload.size());
          if (n)
            memcpy(buf, it->payload.data(), n);
          w.self_box.erase(it);
        }
        return;
      }
      if (w.size == 1 || source == w.rank) {
        w.cv.wait_for(lk, std::chrono::milliseconds(50));
        if (now_s() - t0 > w.timeout_s) {
          fprintf(stderr, "mpi shim: rank %d timed out waiting for tag %d\n", w.rank, tag);
          abort();
        }
        continue;
      }
    }
    if (w.size > 1) {
      RemoteHit hit = scan_inbox(source == w.rank ? -2 : source, tag);
      if (hit.found) {
        if (st) {
          st->MPI_SOURCE = hit.src;
          st->MPI_TAG = tag;
          st->MPI_ERROR = 0;
          st->shim_bytes = hit.bytes;
        }
        if (consume) {
          FILE *f = fopen(hit.path.c_str(), "rb");
          if (!f) {
            fprintf(stderr, "mpi shim: lost message %s\n", hit.path.c_str());
            abort();
          }
          size_t n = std::min(cap, (size_t)hit.bytes);
          if (n && fread(buf, 1, n, f) != n) {
            fprintf(stderr, "mpi shim: short read %s\n", hit.path.c_str());
            abort();
          }
          fclose(f);
          unlink(hit.path.c_str());
        }
        return;
      }
      usleep(200);
      if (now_s() - t0 > w.timeout_s) {
        fprintf(stderr, "mpi shim: rank %d timed out waiting for (src %d, tag %d)\n",
                w.rank, source, tag);
        abort();
      }
    }
  }
}

template <typename T>
inline void reduce_typed(T *acc, const T *in, int n, MPI_Op op) {
  for (int i = 0; i < n; i++) {
    if (op == MPI_SUM)
      acc[i] = acc[i] + in[i];
    else if (op == MPI_MAX)
      acc[i] = std::max(acc[i], in[i]);
    else
      acc[i] = std::min(acc[i], in[i]);
  }
}

inline void reduce_into(void *acc, const void *in, int n, MPI_Datatype t, MPI_Op op) {
  switch (t) {
  case MPI_CHAR:
    reduce_typed((char *)acc, (const char *)in, n, op);
    break;
  case MPI_UNSIGNED_CHAR:
    reduce_typed((unsigned char *)acc, (const unsigned char *)in, n, op);
    break;
  case MPI_INT:
    reduce_typed((int *)acc, (const int *)in, n, op);
    break;
  case MPI_UNSIGNED:
    reduce_typed((unsigned *)acc, (const unsigned *)in, n, op);
    break;
  case MPI_LONG:
    reduce_typed((long *)acc, (const long *)in, n, op);
    break;
  case MPI_UNSIGNED_LONG:
    reduce_typed((unsigned long *)acc, (const unsigned long *)in, n, op);
    break;
  case MPI_FLOAT:
    reduce_typed((float *)acc, (const float *)in, n, op);
    break;
  default:
    reduce_typed((double *)acc, (const double *)in, n, op);
    break;
  }
}

} // namespace nts_mpi_shim

static inline int MPI_Init_thread(int *, char ***, int required, int *provided) {
  *provided = required;
  (void)nts_mpi_shim::world();
  return 0;
}
static inline int MPI_Finalize() { return 0; }
static inline int MPI_Comm_rank(MPI_Comm, int *rank) {
  *rank = nts_mpi_shim::world().rank;
  return 0;
}
static inline int MPI_Comm_size(MPI_Comm, int *size) {
  *size = nts_mpi_shim::world().size;
  return 0;
}
static inline double MPI_Wtime() { return nts_mpi_shim::now_s(); }

static inline int MPI_Send(const void *buf, int count, MPI_Datatype type, int dst,
                           int tag, MPI_Comm) {
  using namespace nts_mpi_shim;
  World &w = world();
  size_t bytes = (size_t)count * type_size(type);
  if (dst == w.rank) {
    LocalMsg m;
    m.tag = tag;
    m.payload.assign((const char *)buf, (const char *)buf + bytes);
    {
      std::lock_guard<std::mutex> lk(w.mu);
      w.self_box.push_back(std::move(m));
    }
    w.cv.notify_all();
  } else {
    send_remote(buf, bytes, dst, tag);
  }
  return 0;
}

static inline int MPI_Probe(int source, int tag, MPI_Comm, MPI_Status *status) {
  MPI_Status local;
  nts_mpi_shim::match(source, tag, false, nullptr, 0, status ? status : &local);
  return 0;
}

static inline int MPI_Get_count(const MPI_Status *status, MPI_Datatype type, int *count) {
  *count = (int)(status->shim_bytes / (long)nts_mpi_shim::type_size(type));
  return 0;
}

static inline int MPI_Recv(void *buf, int count, MPI_Datatype type, int source, int tag,
                           MPI_Comm, MPI_Status *status) {
  nts_mpi_shim::match(source, tag, true, buf,
                      (size_t)count * nts_mpi_shim::type_size(type), status);
  return 0;
}

static inline int MPI_Bcast(void *buf, int count, MPI_Datatype type, int root, MPI_Comm) {
  using namespace nts_mpi_shim;
  World &w = world();
  if (w.size == 1)
    return 0;
  if (w.rank == root) {
    for (int r = 0; r < w.size; r++)
      if (r != root)
        MPI_Send(buf, count, type, r, kCollectiveTag, 0);
  } else {
    MPI_Recv(buf, count, type, root, kCollectiveTag, 0, MPI_STATUS_IGNORE);
  }
  return 0;
}

static inline int MPI_Allreduce(const void *sendbuf, void *recvbuf, int count,
                                MPI_Datatype type, MPI_Op op, MPI_Comm) {
  using namespace nts_mpi_shim;
  World &w = world();
  size_t bytes = (size_t)count * type_size(type);
  if (sendbuf != MPI_IN_PLACE && sendbuf != recvbuf)
    memcpy(recvbuf, sendbuf, bytes);
  if (w.size == 1)
    return 0;
  if (w.rank == 0) {
    std::vector<char> tmp(bytes ? bytes : 1);
    for (int r = 1; r < w.size; r++) {
      MPI_Recv(tmp.data(), count, type, r, kCollectiveTag, 0, MPI_STATUS_IGNORE);
      reduce_into(recvbuf, tmp.data(), count, type, op);
    }
    for (int r = 1; r < w.size; r++)
      MPI_Send(recvbuf, count, type, r, kCollectiveTag, 0);
  } else {
    MPI_Send(recvbuf, count, type, 0, kCollectiveTag, 0);
    MPI_Recv(recvbuf, count, type, 0, kCollectiveTag, 0, MPI_STATUS_IGNORE);
  }
  return 0;
}

static inline int MPI_Barrier(MPI_Comm) {
  int token = 1, out = 0;
  return MPI_Allreduce(&token, &out, 1, MPI_INT, MPI_SUM, 0);
}
