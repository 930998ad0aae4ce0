// minimpi.cpp -- single-node MPI subset over one POSIX shared-memory segment.
// See include/mpi.h for scope.  Every collective is built on one primitive,
// `gather_all`: each member of a communicator publishes <= PAYLOAD bytes in its own
// slot tagged (context id, sequence number); members read each other's slots and
// acknowledge; a slot is rewritten only after all readers of its previous content
// have acknowledged.  Since MPI requires all members of a communicator to call
// collectives on it in the same order, waits always point to strictly earlier
// operations and the protocol cannot deadlock.
#include <mpi.h>

#include <errno.h>
#include <fcntl.h>
#include <sched.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <string>
#include <vector>

namespace {

constexpr uint64_t kMagic = 0x70666674623230ull;  // "pfftb20"
constexpr int kPayload = 4096;
constexpr int kMaxRanks = 64;

struct alignas(128) Slot {
  std::atomic<uint64_t> tag;        // (ctx << 32) | seq of the published payload
  std::atomic<int32_t> readers;     // readers that have not yet consumed it
  int32_t nbytes;
  alignas(16) unsigned char payload[kPayload];
};

struct Segment {
  std::atomic<uint64_t> magic;
  int32_t size;
  std::atomic<int32_t> attached;
  std::atomic<uint32_t> next_ctx;
  std::atomic<int32_t> abort_flag;
  Slot slots[kMaxRanks];
};

struct World {
  bool initialized = false, finalized = false, bootstrapped = false;
  int rank = 0, size = 1;
  std::string shm_name;
  bool owner = false;
  Segment *seg = nullptr;
  Segment local;  // used when size == 1 (no shared memory needed)
};
World g;

double now_s() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

double wait_limit_s() {
  static double lim = [] {
    const char *e = getenv("PFFT_MPI_TIMEOUT");
    return e ? atof(e) : 300.0;
  }();
  return lim;
}

[[noreturn]] void die(const char *what) {
  fprintf(stderr, "minimpi[rank %d/%d]: %s\n", g.rank, g.size, what);
  if (g.seg) g.seg->abort_flag.store(1);
  _exit(86);
}

struct Spinner {
  int n = 0;
  double t0 = 0;
  void pause() {
    if (++n < 2000) {
#if defined(__x86_64__)
      __builtin_ia32_pause();
#endif
      return;
    }
    sched_yield();
    if ((n & 0x3ff) == 0) {
      if (g.seg && g.seg->abort_flag.load(std::memory_order_relaxed)) {
        fprintf(stderr, "minimpi[rank %d]: another rank aborted\n", g.rank);
        _exit(87);
      }
      double t = now_s();
      if (t0 == 0) t0 = t;
      else if (t - t0 > wait_limit_s()) die("timeout waiting for a peer (PFFT_MPI_TIMEOUT)");
    }
  }
};

size_t dt_size(MPI_Datatype t) {
  switch (t) {
    case MPI_CHAR: case MPI_BYTE: return 1;
    case MPI_INT: case MPI_UNSIGNED: return 4;
    case MPI_FLOAT: return 4;
    case MPI_LONG: case MPI_UNSIGNED_LONG: case MPI_LONG_LONG: case MPI_AINT: return 8;
    case MPI_DOUBLE: return 8;
    case MPI_LONG_DOUBLE: return sizeof(long double);
    default: return 0;
  }
}

}  // namespace

struct minimpi_comm_s {
  bool builtin = false;       // WORLD / SELF
  bool is_self = false;
  uint32_t ctx = 0;
  uint32_t seq = 0;
  int rank = 0;
  std::vector<int> members;   // world ranks, indexed by rank in this communicator
  int ndims = 0;              // > 0: Cartesian topology
  int dims[8] = {0}, periods[8] = {0};
};

struct minimpi_comm_s minimpi_comm_world_obj;
struct minimpi_comm_s minimpi_comm_self_obj;

namespace {

void setup_builtin() {
  auto &w = minimpi_comm_world_obj;
  w.builtin = true;
  w.ctx = 1;
  w.rank = g.rank;
  w.members.resize(g.size);
  for (int i = 0; i < g.size; i++) w.members[i] = i;
  auto &s = minimpi_comm_self_obj;
  s.builtin = true;
  s.is_self = true;
  s.ctx = 2;
  s.rank = 0;
  s.members.assign(1, g.rank);
}

void attach_segment() {
  if (g.size == 1) {
    g.seg = &g.local;
    g.seg->size = 1;
    g.seg->next_ctx.store(16);
    return;
  }
  if (g.size > kMaxRanks) die("too many ranks for minimpi (max 64)");
  const char *name = g.shm_name.c_str();
  int fd = -1;
  const char *precreated = getenv("PFFT_MPI_PRECREATED");
  if (g.rank == 0 && !precreated) {
    shm_unlink(name);
    fd = shm_open(name, O_CREAT | O_EXCL | O_RDWR, 0600);
    if (fd < 0) die("shm_open(create) failed");
    if (ftruncate(fd, sizeof(Segment)) != 0) die("ftruncate failed");
    g.owner = true;
  } else {
    Spinner sp;
    while ((fd = shm_open(name, O_RDWR, 0600)) < 0) {
      usleep(2000);
      sp.n = 4000;
      sp.pause();
    }
    struct stat st;
    Spinner sp2;
    while (fstat(fd, &st) == 0 && (size_t)st.st_size < sizeof(Segment)) {
      usleep(1000);
      sp2.n = 4000;
      sp2.pause();
    }
  }
  void *p = mmap(nullptr, sizeof(Segment), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  close(fd);
  if (p == MAP_FAILED) die("mmap failed");
  g.seg = static_cast<Segment *>(p);
  if (g.rank == 0) {
    // a fresh ftruncate'd segment is zero-filled: tags 0, readers 0
    g.seg->size = g.size;
    g.seg->next_ctx.store(16);
    g.seg->attached.store(0);
    g.seg->magic.store(kMagic, std::memory_order_release);
  } else {
    Spinner sp;
    while (g.seg->magic.load(std::memory_order_acquire) != kMagic) sp.pause();
    if (g.seg->size != g.size) die("world size mismatch in shared segment");
  }
  g.seg->attached.fetch_add(1);
  Spinner sp;
  while (g.seg->attached.load() < g.size) sp.pause();
}

// The one primitive: all-gather of `nbytes` (<= kPayload) per member.
// `out` (may be null) receives members*nbytes bytes ordered by communicator rank.
void gather_all(MPI_Comm c, const void *mine, int nbytes, void *out) {
  const int n = (int)c->members.size();
  if (out && nbytes) memcpy((char *)out + (size_t)c->rank * nbytes, mine, nbytes);
  if (n == 1) return;
  const uint64_t tag = ((uint64_t)c->ctx << 32) | (uint64_t)(++c->seq);
  Slot &my = g.seg->slots[c->members[c->rank]];
  {
    Spinner sp;
    while (my.readers.load(std::memory_order_acquire) != 0) sp.pause();
  }
  if (nbytes) memcpy(my.payload, mine, nbytes);
  my.nbytes = nbytes;
  my.readers.store(n - 1, std::memory_order_relaxed);
  my.tag.store(tag, std::memory_order_release);
  for (int k = 1; k < n; k++) {
    int r = (c->rank + k) % n;
    Slot &s = g.seg->slots[c->members[r]];
    Spinner sp;
    while (s.tag.load(std::memory_order_acquire) != tag) sp.pause();
    if (out && nbytes) memcpy((char *)out + (size_t)r * nbytes, s.payload, nbytes);
    s.readers.fetch_sub(1, std::memory_order_release);
  }
}

// all-gather of arbitrary size, chunked
void gather_all_big(MPI_Comm c, const void *mine, size_t nbytes, void *out) {
  const int n = (int)c->members.size();
  if (nbytes <= (size_t)kPayload) {
    gather_all(c, mine, (int)nbytes, out);
    return;
  }
  std::vector<unsigned char> tmp((size_t)kPayload * n);
  for (size_t off = 0; off < nbytes; off += kPayload) {
    int len = (int)std::min<size_t>(kPayload, nbytes - off);
    gather_all(c, (const char *)mine + off, len, tmp.data());
    for (int r = 0; r < n; r++) memcpy((char *)out + (size_t)r * nbytes + off, tmp.data() + (size_t)r * len, len);
  }
}

template <class T>
void reduce_typed(T *acc, const T *x, int count, MPI_Op op) {
  for (int i = 0; i < count; i++) {
    switch (op) {
      case MPI_MAX: acc[i] = std::max(acc[i], x[i]); break;
      case MPI_MIN: acc[i] = std::min(acc[i], x[i]); break;
      case MPI_SUM: acc[i] = acc[i] + x[i]; break;
      case MPI_PROD: acc[i] = acc[i] * x[i]; break;
      case MPI_LAND: acc[i] = (T)((acc[i] != 0) && (x[i] != 0)); break;
      case MPI_LOR: acc[i] = (T)((acc[i] != 0) || (x[i] != 0)); break;
      default: break;
    }
  }
}

void reduce_bytes(void *acc, const void *x, int count, MPI_Datatype t, MPI_Op op) {
  switch (t) {
    case MPI_CHAR: reduce_typed((signed char *)acc, (const signed char *)x, count, op); break;
    case MPI_BYTE: reduce_typed((unsigned char *)acc, (const unsigned char *)x, count, op); break;
    case MPI_INT: reduce_typed((int *)acc, (const int *)x, count, op); break;
    case MPI_UNSIGNED: reduce_typed((unsigned *)acc, (const unsigned *)x, count, op); break;
    case MPI_LONG: case MPI_LONG_LONG: case MPI_AINT:
      reduce_typed((long long *)acc, (const long long *)x, count, op); break;
    case MPI_UNSIGNED_LONG: reduce_typed((unsigned long long *)acc, (const unsigned long long *)x, count, op); break;
    case MPI_FLOAT: reduce_typed((float *)acc, (const float *)x, count, op); break;
    case MPI_DOUBLE: reduce_typed((double *)acc, (const double *)x, count, op); break;
    case MPI_LONG_DOUBLE: reduce_typed((long double *)acc, (const long double *)x, count, op); break;
    default: break;
  }
}

void ensure_init() {
  if (!g.initialized) MPI_Init(nullptr, nullptr);
}

// Build a new communicator out of `parent` members having the same color; collective on parent.
MPI_Comm split_impl(MPI_Comm parent, int color, int key) {
  const int n = (int)parent->members.size();
  struct CK { int color, key; };
  CK mine{color, key};
  std::vector<CK> all(n);
  gather_all(parent, &mine, sizeof(CK), all.data());
  // group leader (lowest parent rank of each color) draws a fresh context id
  std::vector<int> grp;
  for (int r = 0; r < n; r++)
    if (all[r].color == color) grp.push_back(r);
  std::stable_sort(grp.begin(), grp.end(), [&](int a, int b) { return all[a].key < all[b].key; });
  int leader = *std::min_element(grp.begin(), grp.end());
  uint32_t ctx = 0;
  if (color != MPI_UNDEFINED && leader == parent->rank) ctx = g.seg->next_ctx.fetch_add(1);
  std::vector<uint32_t> ctxs(n);
  gather_all(parent, &ctx, sizeof(ctx), ctxs.data());
  if (color == MPI_UNDEFINED) return MPI_COMM_NULL;
  MPI_Comm c = new minimpi_comm_s;
  c->ctx = ctxs[leader];
  for (size_t i = 0; i < grp.size(); i++) {
    c->members.push_back(parent->members[grp[i]]);
    if (grp[i] == parent->rank) c->rank = (int)i;
  }
  return c;
}

}  // namespace

extern "C" {

int minimpi_bootstrap(const char *jobname, int rank, int size) {
  if (g.initialized) return g.size == size && g.rank == rank ? 0 : 1;
  g.bootstrapped = true;
  g.rank = rank;
  g.size = size;
  g.shm_name = std::string("/pfftb200_") + jobname;
  return 0;
}

int minimpi_world_rank(MPI_Comm comm, int r) {
  if (!comm || r < 0 || r >= (int)comm->members.size()) return -1;
  return comm->members[r];
}

int minimpi_is_parallel(void) { return g.size > 1; }

int MPI_Init(int *, char ***) {
  if (g.initialized) return MPI_SUCCESS;
  if (!g.bootstrapped) {
    const char *job = getenv("PFFT_MPI_JOB");
    const char *r = getenv("PFFT_MPI_RANK"), *s = getenv("PFFT_MPI_SIZE");
    if (job && r && s) {
      g.rank = atoi(r);
      g.size = atoi(s);
      g.shm_name = std::string("/pfftb200_") + job;
    } else if (getenv("RANK") && getenv("WORLD_SIZE") && atoi(getenv("WORLD_SIZE")) > 1) {
      // torchrun-style environment: all workers share parent pid, port and run id
      g.rank = atoi(getenv("RANK"));
      g.size = atoi(getenv("WORLD_SIZE"));
      const char *port = getenv("MASTER_PORT");
      const char *rid = getenv("TORCHELASTIC_RUN_ID");
      char buf[256];
      snprintf(buf, sizeof buf, "/pfftb200_tr_%s_%s_%d", port ? port : "0", rid ? rid : "none", (int)getppid());
      for (char *p = buf + 1; *p; p++)
        if (*p == '/') *p = '_';
      g.shm_name = buf;
    }
  }
  attach_segment();
  setup_builtin();
  g.initialized = true;
  return MPI_SUCCESS;
}

int MPI_Init_thread(int *argc, char ***argv, int required, int *provided) {
  if (provided) *provided = required;
  return MPI_Init(argc, argv);
}

int MPI_Initialized(int *flag) { *flag = g.initialized; return MPI_SUCCESS; }
int MPI_Finalized(int *flag) { *flag = g.finalized; return MPI_SUCCESS; }

int MPI_Finalize(void) {
  if (!g.initialized || g.finalized) return MPI_SUCCESS;
  gather_all(MPI_COMM_WORLD, nullptr, 0, nullptr);
  if (g.size > 1) {
    if (g.owner) shm_unlink(g.shm_name.c_str());
    munmap(g.seg, sizeof(Segment));
    g.seg = nullptr;
  }
  g.finalized = true;
  return MPI_SUCCESS;
}

int MPI_Abort(MPI_Comm, int errorcode) {
  if (g.seg) g.seg->abort_flag.store(1);
  if (g.owner) shm_unlink(g.shm_name.c_str());
  _exit(errorcode ? errorcode : 1);
}

double MPI_Wtime(void) { return now_s(); }
double MPI_Wtick(void) { return 1e-9; }

int MPI_Get_processor_name(char *name, int *resultlen) {
  if (gethostname(name, MPI_MAX_PROCESSOR_NAME) != 0) strcpy(name, "localhost");
  *resultlen = (int)strlen(name);
  return MPI_SUCCESS;
}

int MPI_Comm_size(MPI_Comm c, int *size) {
  ensure_init();
  if (!c) return MPI_ERR_COMM;
  *size = (int)c->members.size();
  return MPI_SUCCESS;
}

int MPI_Comm_rank(MPI_Comm c, int *rank) {
  ensure_init();
  if (!c) return MPI_ERR_COMM;
  *rank = c->rank;
  return MPI_SUCCESS;
}

int MPI_Comm_dup(MPI_Comm c, MPI_Comm *newcomm) {
  ensure_init();
  if (!c) return MPI_ERR_COMM;
  MPI_Comm d = split_impl(c, 0, c->rank);
  d->ndims = c->ndims;
  memcpy(d->dims, c->dims, sizeof d->dims);
  memcpy(d->periods, c->periods, sizeof d->periods);
  d->is_self = c->is_self;
  *newcomm = d;
  return MPI_SUCCESS;
}

int MPI_Comm_split(MPI_Comm c, int color, int key, MPI_Comm *newcomm) {
  ensure_init();
  if (!c) return MPI_ERR_COMM;
  *newcomm = split_impl(c, color, key);
  return MPI_SUCCESS;
}

int MPI_Comm_free(MPI_Comm *c) {
  if (!c || !*c) return MPI_ERR_COMM;
  if (!(*c)->builtin) delete *c;
  *c = MPI_COMM_NULL;
  return MPI_SUCCESS;
}

int MPI_Topo_test(MPI_Comm c, int *status) {
  if (!c) return MPI_ERR_COMM;
  *status = c->ndims > 0 ? MPI_CART : MPI_UNDEFINED;
  return MPI_SUCCESS;
}

int MPI_Cart_create(MPI_Comm old, int ndims, const int *dims, const int *periods, int, MPI_Comm *cart) {
  ensure_init();
  if (!old) return MPI_ERR_COMM;
  if (ndims < 1 || ndims > 8) return MPI_ERR_ARG;
  long prod = 1;
  for (int t = 0; t < ndims; t++) prod *= dims[t];
  if (prod > (long)old->members.size()) return MPI_ERR_ARG;
  int in = old->rank < prod;
  MPI_Comm c = split_impl(old, in ? 0 : MPI_UNDEFINED, old->rank);
  if (c) {
    c->ndims = ndims;
    for (int t = 0; t < ndims; t++) {
      c->dims[t] = dims[t];
      c->periods[t] = periods ? periods[t] : 0;
    }
  }
  *cart = c;
  return MPI_SUCCESS;
}

int MPI_Cartdim_get(MPI_Comm c, int *ndims) {
  if (!c || c->ndims == 0) return MPI_ERR_COMM;
  *ndims = c->ndims;
  return MPI_SUCCESS;
}

int MPI_Cart_coords(MPI_Comm c, int rank, int maxdims, int *coords) {
  if (!c || c->ndims == 0) return MPI_ERR_COMM;
  for (int t = c->ndims - 1; t >= 0; t--) {
    if (t < maxdims) coords[t] = rank % c->dims[t];
    rank /= c->dims[t];
  }
  return MPI_SUCCESS;
}

int MPI_Cart_get(MPI_Comm c, int maxdims, int *dims, int *periods, int *coords) {
  if (!c || c->ndims == 0) return MPI_ERR_COMM;
  for (int t = 0; t < c->ndims && t < maxdims; t++) {
    dims[t] = c->dims[t];
    periods[t] = c->periods[t];
  }
  return MPI_Cart_coords(c, c->rank, maxdims, coords);
}

int MPI_Cart_rank(MPI_Comm c, const int *coords, int *rank) {
  if (!c || c->ndims == 0) return MPI_ERR_COMM;
  int r = 0;
  for (int t = 0; t < c->ndims; t++) {
    int x = coords[t];
    if (c->periods[t]) x = ((x % c->dims[t]) + c->dims[t]) % c->dims[t];
    if (x < 0 || x >= c->dims[t]) return MPI_ERR_ARG;
    r = r * c->dims[t] + x;
  }
  *rank = r;
  return MPI_SUCCESS;
}

int MPI_Cart_shift(MPI_Comm c, int direction, int disp, int *src, int *dst) {
  if (!c || c->ndims == 0 || direction < 0 || direction >= c->ndims) return MPI_ERR_COMM;
  int coords[8];
  MPI_Cart_coords(c, c->rank, 8, coords);
  auto at = [&](int delta) {
    int cc[8];
    memcpy(cc, coords, sizeof cc);
    cc[direction] += delta;
    int r;
    if (!c->periods[direction] && (cc[direction] < 0 || cc[direction] >= c->dims[direction])) return MPI_PROC_NULL;
    MPI_Cart_rank(c, cc, &r);
    return r;
  };
  *src = at(-disp);
  *dst = at(+disp);
  return MPI_SUCCESS;
}

int MPI_Cart_sub(MPI_Comm c, const int *remain, MPI_Comm *newcomm) {
  ensure_init();
  if (!c || c->ndims == 0) return MPI_ERR_COMM;
  int coords[8];
  MPI_Cart_coords(c, c->rank, 8, coords);
  int color = 0, key = 0;
  for (int t = 0; t < c->ndims; t++) {
    if (remain[t]) key = key * c->dims[t] + coords[t];
    else color = color * c->dims[t] + coords[t];
  }
  MPI_Comm s = split_impl(c, color, key);
  int k = 0;
  for (int t = 0; t < c->ndims; t++)
    if (remain[t]) {
      s->dims[k] = c->dims[t];
      s->periods[k] = c->periods[t];
      k++;
    }
  s->ndims = k;
  if (k == 0) {  // MPI: zero-dimensional Cartesian communicator of one process
    s->ndims = 1;
    s->dims[0] = 1;
    s->periods[0] = 0;
  }
  *newcomm = s;
  return MPI_SUCCESS;
}

int MPI_Dims_create(int nnodes, int ndims, int *dims) {
  // balanced factorisation, largest first (only free entries == 0 are filled)
  int fixed = 1, nfree = 0;
  for (int t = 0; t < ndims; t++) {
    if (dims[t] > 0) fixed *= dims[t];
    else nfree++;
  }
  if (nfree == 0) return MPI_SUCCESS;
  int rest = nnodes / fixed;
  std::vector<int> f(nfree, 1);
  for (int p = 2; rest > 1;) {
    if (rest % p == 0) {
      *std::min_element(f.begin(), f.end()) *= p;
      rest /= p;
    } else p++;
  }
  std::sort(f.begin(), f.end(), [](int a, int b) { return a > b; });
  int k = 0;
  for (int t = 0; t < ndims; t++)
    if (dims[t] <= 0) dims[t] = f[k++];
  return MPI_SUCCESS;
}

int MPI_Barrier(MPI_Comm c) {
  ensure_init();
  if (!c) return MPI_ERR_COMM;
  gather_all(c, nullptr, 0, nullptr);
  return MPI_SUCCESS;
}

int MPI_Allgather(const void *sb, int scount, MPI_Datatype st, void *rb, int, MPI_Datatype, MPI_Comm c) {
  ensure_init();
  if (!c) return MPI_ERR_COMM;
  size_t nb = dt_size(st) * (size_t)scount;
  std::vector<unsigned char> mine(nb ? nb : 1);
  if (sb == MPI_IN_PLACE) memcpy(mine.data(), (char *)rb + nb * c->rank, nb);
  else memcpy(mine.data(), sb, nb);
  gather_all_big(c, mine.data(), nb, rb);
  return MPI_SUCCESS;
}

int MPI_Gather(const void *sb, int scount, MPI_Datatype st, void *rb, int rc, MPI_Datatype rt, int root, MPI_Comm c) {
  ensure_init();
  if (!c) return MPI_ERR_COMM;
  size_t nb = dt_size(st) * (size_t)scount;
  std::vector<unsigned char> all(nb * c->members.size() + 1);
  gather_all_big(c, sb, nb, all.data());
  if (c->rank == root) memcpy(rb, all.data(), nb * c->members.size());
  (void)rc; (void)rt;
  return MPI_SUCCESS;
}

int MPI_Bcast(void *buf, int count, MPI_Datatype t, int root, MPI_Comm c) {
  ensure_init();
  if (!c) return MPI_ERR_COMM;
  size_t nb = dt_size(t) * (size_t)count;
  const int n = (int)c->members.size();
  // payloads are tiny (ids, handles, sizes): an all-gather from which everybody keeps root's part
  std::vector<unsigned char> all(nb * n + 1);
  gather_all_big(c, buf, nb, all.data());
  if (c->rank != root) memcpy(buf, all.data() + nb * root, nb);
  return MPI_SUCCESS;
}

int MPI_Allreduce(const void *sb, void *rb, int count, MPI_Datatype t, MPI_Op op, MPI_Comm c) {
  ensure_init();
  if (!c) return MPI_ERR_COMM;
  size_t es = dt_size(t);
  if (!es) return MPI_ERR_ARG;
  size_t nb = es * (size_t)count;
  const int n = (int)c->members.size();
  std::vector<unsigned char> mine(nb + 1), all(nb * n + 1);
  memcpy(mine.data(), sb == MPI_IN_PLACE ? rb : sb, nb);
  gather_all_big(c, mine.data(), nb, all.data());
  memcpy(rb, all.data(), nb);
  for (int r = 1; r < n; r++) reduce_bytes(rb, all.data() + nb * r, count, t, op);
  return MPI_SUCCESS;
}

int MPI_Reduce(const void *sb, void *rb, int count, MPI_Datatype t, MPI_Op op, int root, MPI_Comm c) {
  ensure_init();
  if (!c) return MPI_ERR_COMM;
  size_t nb = dt_size(t) * (size_t)count;
  std::vector<unsigned char> res(nb + 1);
  const void *src = sb == MPI_IN_PLACE ? rb : sb;
  int rc = MPI_Allreduce(src, res.data(), count, t, op, c);
  if (c->rank == root) memcpy(rb, res.data(), nb);
  return rc;
}

}  // extern "C"
