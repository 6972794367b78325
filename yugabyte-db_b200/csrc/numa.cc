// numa.cc — host threads and staging memory next to the GPU they feed.
//
// The end-to-end path is bound by host<->device DMA (DESIGN.md "End-to-end modes"). On a two-socket
// box a copy whose pinned buffer lives on the other socket crosses the inter-socket link on top of
// PCIe, and several GPUs doing so share that link. ybgpu_bind_thread_to_device() pins the CALLING thread
// to the CPUs of the NUMA node the device hangs off (sysfs: /sys/bus/pci/devices/<bdf>/numa_node,
// /sys/devices/system/node/node<N>/cpulist) and makes that node the thread's preferred memory node
// (set_mempolicy(MPOL_PREFERRED)); threads created afterwards inherit both, and first-touch / pinned
// allocations made by them land next to the GPU. No libnuma: two syscalls and sysfs.
//
// Reference context: the reference runs compactions on a yb::PriorityThreadPool whose threads are not
// bound (rocksdb/db/db_impl.cc:397-403); binding is what a device-attached worker needs in addition.
#include <cuda_runtime.h>
#include <sched.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "../../include/ybgpu_compaction.h"

namespace {

bool ReadFile(const std::string& path, std::string* out) {
  FILE* f = fopen(path.c_str(), "r");
  if (!f) return false;
  char buf[4096];
  size_t n = fread(buf, 1, sizeof(buf) - 1, f);
  fclose(f);
  buf[n] = 0;
  *out = buf;
  return true;
}

// "0-31,64-95" -> cpu_set_t
int ParseCpuList(const std::string& s, cpu_set_t* set) {
  CPU_ZERO(set);
  int count = 0;
  const char* p = s.c_str();
  while (*p) {
    while (*p && !isdigit(static_cast<unsigned char>(*p))) p++;
    if (!*p) break;
    char* e;
    long a = strtol(p, &e, 10), b = a;
    p = e;
    if (*p == '-') { b = strtol(p + 1, &e, 10); p = e; }
    for (long c = a; c <= b && c < CPU_SETSIZE; c++) { CPU_SET(static_cast<int>(c), set); count++; }
  }
  return count;
}

}  // namespace

extern "C" {

int32_t ybgpu_device_numa_node(int32_t device) {
  char bdf[64] = {0};
  if (cudaDeviceGetPCIBusId(bdf, sizeof(bdf), device) != cudaSuccess) return -1;
  for (char* c = bdf; *c; c++) *c = static_cast<char>(tolower(static_cast<unsigned char>(*c)));
  std::string txt;
  if (!ReadFile(std::string("/sys/bus/pci/devices/") + bdf + "/numa_node", &txt)) return -1;
  return static_cast<int32_t>(atoi(txt.c_str()));
}

ybgpu_status ybgpu_bind_thread_to_device(int32_t device, int32_t* numa_node, int32_t* num_cpus) {
  if (numa_node) *numa_node = -1;
  if (num_cpus) *num_cpus = 0;
  const char* env = getenv("YBGPU_NUMA_BIND");
  if (env && atoi(env) == 0) return YBGPU_OK;
  const int32_t node = ybgpu_device_numa_node(device);
  if (node < 0) return YBGPU_OK;                         // single-node host or no sysfs: nothing to do
  std::string cpus;
  if (!ReadFile("/sys/devices/system/node/node" + std::to_string(node) + "/cpulist", &cpus)) return YBGPU_OK;
  cpu_set_t set;
  const int n = ParseCpuList(cpus, &set);
  if (n == 0) return YBGPU_OK;
  // only CPUs this process may use (cgroup / taskset limits stay in force)
  cpu_set_t allowed;
  if (sched_getaffinity(0, sizeof(allowed), &allowed) == 0) {
    cpu_set_t both;
    CPU_AND(&both, &set, &allowed);
    if (CPU_COUNT(&both) == 0) return YBGPU_OK;
    set = both;
  }
  if (sched_setaffinity(0, sizeof(set), &set) != 0) return YBGPU_OK;
  if (node < 1024) {
    unsigned long mask[16] = {0};
    mask[node / (8 * sizeof(unsigned long))] |= 1ul << (node % (8 * sizeof(unsigned long)));
    const int MPOL_PREFERRED_ = 1;
    syscall(SYS_set_mempolicy, MPOL_PREFERRED_, mask, sizeof(mask) * 8);     // best effort
  }
  if (numa_node) *numa_node = node;
  if (num_cpus) *num_cpus = CPU_COUNT(&set);
  return YBGPU_OK;
}

}  // extern "C"
