// definitions of the emulated CUDA built-ins (see cuda_runtime.h in this directory)
#include "cuda_runtime.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <map>
#include <mutex>
#include <string>

thread_local emu_uint3 threadIdx;
emu_uint3 blockIdx, blockDim, gridDim;
pthread_barrier_t emu_block_barrier;
namespace p3b { alignas(16) unsigned char smem_raw[232448]; }

namespace {
struct Area {
  std::string name;
  size_t bytes;
  bool owner;
};
std::mutex g_mu;
std::map<void *, Area> g_areas;
int g_counter = 0;
struct Handle {  // what travels in the 64 IPC bytes
  char name[48];
  unsigned long long bytes;
};
static_assert(sizeof(Handle) <= sizeof(cudaIpcMemHandle_t), "handle size");

void *map_shm(const std::string &name, size_t bytes, bool create) {
  int fd = shm_open(name.c_str(), create ? (O_CREAT | O_EXCL | O_RDWR) : O_RDWR, 0600);
  if (fd < 0) return nullptr;
  if (create && ftruncate(fd, (off_t)bytes) != 0) {
    close(fd);
    shm_unlink(name.c_str());
    return nullptr;
  }
  void *p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  close(fd);
  return p == MAP_FAILED ? nullptr : p;
}
}  // namespace

cudaError_t cudaMalloc(void **p, size_t n) {
  std::lock_guard<std::mutex> lk(g_mu);
  size_t bytes = (n + 4095) & ~size_t(4095);
  if (!bytes) bytes = 4096;
  char nm[48];
  snprintf(nm, sizeof nm, "/p3b_emu_%d_%d", (int)getpid(), g_counter++);
  void *q = map_shm(nm, bytes, true);
  if (!q) return 2;
  g_areas[q] = Area{nm, bytes, true};
  *p = q;
  return cudaSuccess;
}

cudaError_t cudaFree(void *p) {
  if (!p) return cudaSuccess;
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_areas.find(p);
  if (it == g_areas.end()) return 1;
  munmap(p, it->second.bytes);
  if (it->second.owner) shm_unlink(it->second.name.c_str());
  g_areas.erase(it);
  return cudaSuccess;
}

cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *h, void *p) {
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_areas.find(p);
  if (it == g_areas.end() || !it->second.owner) return 1;
  Handle hd;
  memset(&hd, 0, sizeof hd);
  strncpy(hd.name, it->second.name.c_str(), sizeof hd.name - 1);
  hd.bytes = it->second.bytes;
  memset(h, 0, sizeof *h);
  memcpy(h, &hd, sizeof hd);
  return cudaSuccess;
}

cudaError_t cudaIpcOpenMemHandle(void **p, cudaIpcMemHandle_t h, int) {
  Handle hd;
  memcpy(&hd, &h, sizeof hd);
  void *q = map_shm(hd.name, (size_t)hd.bytes, false);
  if (!q) return 1;
  std::lock_guard<std::mutex> lk(g_mu);
  g_areas[q] = Area{hd.name, (size_t)hd.bytes, false};
  *p = q;
  return cudaSuccess;
}

cudaError_t cudaIpcCloseMemHandle(void *p) { return cudaFree(p); }

// free every shared-memory file this process still owns (atexit)
namespace {
struct Reaper {
  ~Reaper() {
    for (auto &kv : g_areas)
      if (kv.second.owner) shm_unlink(kv.second.name.c_str());
  }
} g_reaper;
}  // namespace
