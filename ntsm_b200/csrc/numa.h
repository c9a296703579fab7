// numa.h -- keep a GPU's pinned batch ring and the threads that fill it on the GPU's NUMA node
// ("NUMA-aware pinned rings").  On a two-socket host a batch packed by a core of one
// socket into memory of the other, then DMA'd by a GPU behind the first, crosses the socket
// interconnect twice.  Everything here is best effort and silent: a box with one node (the VMs this was
// measured on expose exactly one, and report numa_node = -1 for every GPU: profiles/r02_box_topology.txt)
// or without the sysfs files behaves exactly as before.
#pragma once
#include <string>
#include <vector>

namespace ntsm {

// NUMA node of CUDA device `device` from /sys/bus/pci/devices/<bdf>/numa_node; -1 = unknown / single node
int gpu_numa_node(int device);
// "0-3,8,10-11" -> {0,1,2,3,8,10,11}  (the format of /sys/devices/system/node/nodeN/cpulist)
std::vector<int> parse_cpulist(const std::string &text);
int numa_node_count();
// cpus of `node` that this process may run on (empty when unknown)
std::vector<int> node_cpus(int node);

// While alive, new pages of the calling thread are preferably taken from `node` (set_mempolicy
// MPOL_PREFERRED) -- cudaMallocHost's pages are allocated and touched by the calling thread.  node < 0: no-op.
struct PreferNode {
	explicit PreferNode(int node);
	~PreferNode();
	bool active = false;
};
// Pin the calling thread to the cpus of `node` (no-op when node < 0 or unknown); returns true if it did.
bool run_on_node(int node);

}  // namespace ntsm
