// mpi_shim.cpp -- implementation of the threads-as-ranks MPI substitute (see mpi.h).
// TEST INFRASTRUCTURE ONLY: used to build the unmodified reference CPU solver as
// the parity oracle / CPU baseline.  Nothing in the product path links this.
#include "mpi.h"

#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <thread>
#include <unordered_map>
#include <vector>

#include <fcntl.h>
#include <unistd.h>

namespace {

struct IndexedType {
    int base;                     // MPI_DOUBLE / MPI_FLOAT / MPI_INT
    std::vector<int> blocklens;
    std::vector<int> displs;
    size_t packed_elems;
};

struct Mailbox {
    std::mutex m;
    std::deque<std::vector<char>> q;
    std::atomic<int> pending{0};
};

struct World {
    int n = 1;
    std::atomic<int> arrived{0};
    std::atomic<int> sense{0};
    // reduction scratch: two alternating sets of per-rank slots
    static constexpr int kSlotBytes = 512;
    std::vector<char> slots[2];
    std::vector<Mailbox> boxes;   // [dst * n + src]
    std::mutex type_mutex;
    std::vector<IndexedType> types;                        // handle = 16 + index
    std::unordered_map<uint64_t, std::vector<int>> type_index;  // hash -> candidate indices
};

World* g_world = nullptr;
thread_local int t_rank = 0;
thread_local int t_sense = 0;
thread_local int t_slot_set = 0;

inline void cpu_relax(int& spins) {
    if (++spins < 4096) {
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
    } else {
        std::this_thread::yield();
    }
}

void barrier() {
    World& w = *g_world;
    if (w.n == 1) return;
    t_sense ^= 1;
    if (w.arrived.fetch_add(1, std::memory_order_acq_rel) + 1 == w.n) {
        w.arrived.store(0, std::memory_order_relaxed);
        w.sense.store(t_sense, std::memory_order_release);
    } else {
        int spins = 0;
        while (w.sense.load(std::memory_order_acquire) != t_sense) cpu_relax(spins);
    }
}

size_t base_size(MPI_Datatype t) {
    switch (t) {
        case MPI_DOUBLE: return 8;
        case MPI_FLOAT: return 4;
        case MPI_INT: return 4;
        default: return 0;
    }
}

const IndexedType* indexed(MPI_Datatype t) {
    if (t < 16) return nullptr;
    return &g_world->types[static_cast<size_t>(t - 16)];
}

size_t packed_bytes(int count, MPI_Datatype t) {
    if (const IndexedType* it = indexed(t)) return static_cast<size_t>(count) * it->packed_elems * base_size(it->base);
    return static_cast<size_t>(count) * base_size(t);
}

void pack(std::vector<char>& out, const void* buf, int count, MPI_Datatype t) {
    out.resize(packed_bytes(count, t));
    const IndexedType* it = indexed(t);
    if (!it) {
        std::memcpy(out.data(), buf, out.size());
        return;
    }
    const size_t es = base_size(it->base);
    const char* src = static_cast<const char*>(buf);
    char* dst = out.data();
    for (int c = 0; c < count; c++) {
        for (size_t b = 0; b < it->blocklens.size(); b++) {
            const size_t nbytes = static_cast<size_t>(it->blocklens[b]) * es;
            std::memcpy(dst, src + static_cast<size_t>(it->displs[b]) * es, nbytes);
            dst += nbytes;
        }
    }
}

void unpack(const std::vector<char>& in, void* buf, int count, MPI_Datatype t) {
    const IndexedType* it = indexed(t);
    if (!it) {
        std::memcpy(buf, in.data(), in.size());
        return;
    }
    const size_t es = base_size(it->base);
    char* dst = static_cast<char*>(buf);
    const char* src = in.data();
    for (int c = 0; c < count; c++) {
        for (size_t b = 0; b < it->blocklens.size(); b++) {
            const size_t nbytes = static_cast<size_t>(it->blocklens[b]) * es;
            std::memcpy(dst + static_cast<size_t>(it->displs[b]) * es, src, nbytes);
            src += nbytes;
        }
    }
}

template <typename T>
void sum_slots(void* recv, int count, int set) {
    World& w = *g_world;
    T* out = static_cast<T*>(recv);
    for (int i = 0; i < count; i++) {
        T acc = 0;
        for (int r = 0; r < w.n; r++) {  // rank order: deterministic, same on every rank
            const T* s = reinterpret_cast<const T*>(w.slots[set].data() + static_cast<size_t>(r) * World::kSlotBytes);
            acc += s[i];
        }
        out[i] = acc;
    }
}

int reduce_impl(const void* sendbuf, void* recvbuf, int count, MPI_Datatype type, int root) {
    World& w = *g_world;
    const size_t es = base_size(type);
    if (es == 0 || es * static_cast<size_t>(count) > static_cast<size_t>(World::kSlotBytes)) {
        std::fprintf(stderr, "mpi_shim: unsupported reduction (type %d count %d)\n", type, count);
        std::abort();
    }
    if (w.n == 1) {
        if (recvbuf != sendbuf) std::memcpy(recvbuf, sendbuf, es * count);
        return MPI_SUCCESS;
    }
    const int set = t_slot_set;
    t_slot_set ^= 1;
    std::memcpy(w.slots[set].data() + static_cast<size_t>(t_rank) * World::kSlotBytes, sendbuf, es * count);
    barrier();
    if (root < 0 || root == t_rank) {
        if (type == MPI_DOUBLE) sum_slots<double>(recvbuf, count, set);
        else if (type == MPI_FLOAT) sum_slots<float>(recvbuf, count, set);
        else sum_slots<int>(recvbuf, count, set);
    }
    return MPI_SUCCESS;
}

}  // namespace

extern "C" {

int pps_shim_rank(void) { return t_rank; }
int pps_shim_world(void) { return g_world ? g_world->n : 1; }

int pps_shim_run(int world, int (*fn)(int, char**), int argc, char** argv) {
    World w;
    w.n = world;
    w.slots[0].assign(static_cast<size_t>(world) * World::kSlotBytes, 0);
    w.slots[1].assign(static_cast<size_t>(world) * World::kSlotBytes, 0);
    w.boxes = std::vector<Mailbox>(static_cast<size_t>(world) * world);
    g_world = &w;
    std::vector<int> rc(world, 0);
    if (world == 1) {
        t_rank = 0;
        rc[0] = fn(argc, argv);
    } else {
        std::vector<std::thread> threads;
        for (int r = 0; r < world; r++) {
            threads.emplace_back([&, r]() {
                t_rank = r;
                t_sense = 0;
                t_slot_set = 0;
                rc[r] = fn(argc, argv);
            });
        }
        for (auto& t : threads) t.join();
    }
    g_world = nullptr;
    for (int r = 0; r < world; r++) if (rc[r]) return rc[r];
    return 0;
}

int MPI_Init(int*, char***) { return MPI_SUCCESS; }
int MPI_Finalize(void) { return MPI_SUCCESS; }
int MPI_Comm_size(MPI_Comm, int* size) { *size = pps_shim_world(); return MPI_SUCCESS; }
int MPI_Comm_rank(MPI_Comm, int* rank) { *rank = t_rank; return MPI_SUCCESS; }
int MPI_Barrier(MPI_Comm) { barrier(); return MPI_SUCCESS; }

int MPI_Allreduce(const void* sendbuf, void* recvbuf, int count, MPI_Datatype type, MPI_Op, MPI_Comm) {
    return reduce_impl(sendbuf, recvbuf, count, type, -1);
}

int MPI_Reduce(const void* sendbuf, void* recvbuf, int count, MPI_Datatype type, MPI_Op, int root, MPI_Comm) {
    return reduce_impl(sendbuf, recvbuf, count, type, root);
}

int MPI_Bcast(void* buf, int count, MPI_Datatype type, int root, MPI_Comm) {
    World& w = *g_world;
    if (w.n == 1) return MPI_SUCCESS;
    const size_t nbytes = base_size(type) * static_cast<size_t>(count);
    if (nbytes > static_cast<size_t>(World::kSlotBytes)) std::abort();
    const int set = t_slot_set;
    t_slot_set ^= 1;
    if (t_rank == root) std::memcpy(w.slots[set].data(), buf, nbytes);
    barrier();
    if (t_rank != root) std::memcpy(buf, w.slots[set].data(), nbytes);
    return MPI_SUCCESS;
}

int MPI_Isend(const void* buf, int count, MPI_Datatype type, int dest, int, MPI_Comm, MPI_Request* req) {
    World& w = *g_world;
    std::vector<char> msg;
    pack(msg, buf, count, type);
    Mailbox& box = w.boxes[static_cast<size_t>(dest) * w.n + t_rank];
    {
        std::lock_guard<std::mutex> lk(box.m);
        box.q.emplace_back(std::move(msg));
    }
    box.pending.fetch_add(1, std::memory_order_release);
    if (req) { req->kind = 0; req->buf = nullptr; req->count = 0; req->type = 0; req->peer = dest; }
    return MPI_SUCCESS;
}

int MPI_Irecv(void* buf, int count, MPI_Datatype type, int source, int, MPI_Comm, MPI_Request* req) {
    req->kind = 1;
    req->buf = buf;
    req->count = count;
    req->type = type;
    req->peer = source;
    return MPI_SUCCESS;
}

int MPI_Waitall(int count, MPI_Request* reqs, MPI_Status* statuses) {
    World& w = *g_world;
    for (int i = 0; i < count; i++) {
        MPI_Request& r = reqs[i];
        if (r.kind == 1) {
            Mailbox& box = w.boxes[static_cast<size_t>(t_rank) * w.n + r.peer];
            int spins = 0;
            while (box.pending.load(std::memory_order_acquire) == 0) cpu_relax(spins);
            std::vector<char> msg;
            {
                std::lock_guard<std::mutex> lk(box.m);
                msg = std::move(box.q.front());
                box.q.pop_front();
            }
            box.pending.fetch_sub(1, std::memory_order_acq_rel);
            if (msg.size() != packed_bytes(r.count, r.type)) {
                std::fprintf(stderr, "mpi_shim: size mismatch on recv (rank %d from %d: got %zu want %zu)\n",
                             t_rank, r.peer, msg.size(), packed_bytes(r.count, r.type));
                std::abort();
            }
            unpack(msg, r.buf, r.count, r.type);
            r.kind = 0;
        }
        if (statuses) { statuses[i].MPI_SOURCE = r.peer; statuses[i].MPI_TAG = 0; statuses[i].MPI_ERROR = MPI_SUCCESS; }
    }
    return MPI_SUCCESS;
}

int MPI_Type_indexed(int count, const int* blocklens, const int* displs, MPI_Datatype oldtype, MPI_Datatype* newtype) {
    World& w = *g_world;
    // The reference creates (and leaks) two of these per face per exchange
    // (communicationMPI.hpp:98-103): intern identical descriptions.
    uint64_t h = 1469598103934665603ull ^ static_cast<uint64_t>(oldtype);
    auto mix = [&h](uint64_t v) { h ^= v; h *= 1099511628211ull; };
    mix(static_cast<uint64_t>(count));
    for (int i = 0; i < count; i++) { mix(static_cast<uint32_t>(blocklens[i])); mix(static_cast<uint32_t>(displs[i])); }
    std::lock_guard<std::mutex> lk(w.type_mutex);
    auto& cands = w.type_index[h];
    for (int idx : cands) {
        const IndexedType& t = w.types[idx];
        if (t.base == oldtype && static_cast<int>(t.blocklens.size()) == count &&
            std::memcmp(t.blocklens.data(), blocklens, sizeof(int) * count) == 0 &&
            std::memcmp(t.displs.data(), displs, sizeof(int) * count) == 0) {
            *newtype = 16 + idx;
            return MPI_SUCCESS;
        }
    }
    IndexedType t;
    t.base = oldtype;
    t.blocklens.assign(blocklens, blocklens + count);
    t.displs.assign(displs, displs + count);
    t.packed_elems = 0;
    for (int i = 0; i < count; i++) t.packed_elems += static_cast<size_t>(blocklens[i]);
    // keep element addresses stable for readers that hold no lock: reserve generously
    if (w.types.capacity() == 0) w.types.reserve(4096);
    if (w.types.size() == w.types.capacity()) {
        std::fprintf(stderr, "mpi_shim: too many distinct indexed types\n");
        std::abort();
    }
    w.types.push_back(std::move(t));
    const int idx = static_cast<int>(w.types.size()) - 1;
    cands.push_back(idx);
    *newtype = 16 + idx;
    return MPI_SUCCESS;
}

int MPI_Type_commit(MPI_Datatype*) { return MPI_SUCCESS; }
int MPI_Type_free(MPI_Datatype* t) { if (t) *t = MPI_DATATYPE_NULL; return MPI_SUCCESS; }

// MPI-IO subset (alpaka tree, src/main.cpp:137-145): one descriptor + individual file pointer per rank-thread.
struct pps_shim_file {
    int fd;
    long long pos;
};

int MPI_File_open(MPI_Comm, const char* filename, int amode, MPI_Info, MPI_File* fh) {
    int flags = (amode & MPI_MODE_WRONLY) ? O_WRONLY : O_RDONLY;
    if (amode & MPI_MODE_CREATE) flags |= O_CREAT;
    const int fd = ::open(filename, flags, 0644);
    if (fd < 0) { *fh = nullptr; return 1; }
    *fh = new pps_shim_file{fd, 0};
    barrier();  // collective in MPI: nobody writes before everybody has the file
    return MPI_SUCCESS;
}
int MPI_File_seek(MPI_File fh, MPI_Offset offset, int whence) {
    if (!fh || whence != MPI_SEEK_SET) return 1;
    fh->pos = offset;
    return MPI_SUCCESS;
}
int MPI_File_write(MPI_File fh, const void* buf, int count, MPI_Datatype type, MPI_Status*) {
    if (!fh) return 1;
    const size_t nbytes = base_size(type) * static_cast<size_t>(count);
    const char* p = static_cast<const char*>(buf);
    size_t done = 0;
    while (done < nbytes) {
        const ssize_t n = ::pwrite(fh->fd, p + done, nbytes - done, static_cast<off_t>(fh->pos + static_cast<long long>(done)));
        if (n <= 0) return 1;
        done += static_cast<size_t>(n);
    }
    fh->pos += static_cast<long long>(nbytes);
    return MPI_SUCCESS;
}
int MPI_File_close(MPI_File* fh) {
    if (fh && *fh) { ::close((*fh)->fd); delete *fh; *fh = nullptr; }
    barrier();
    return MPI_SUCCESS;
}

}  // extern "C"
