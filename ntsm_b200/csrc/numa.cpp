// numa.cpp -- see numa.h.  No libnuma in the image: raw sysfs reads, sched_setaffinity and the set_mempolicy syscall.
#include "numa.h"

#include <cuda_runtime.h>
#include <sched.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <fstream>

namespace ntsm {

static std::string slurp(const std::string &path)
{
	std::ifstream f(path.c_str());
	std::string s;
	if (f) std::getline(f, s);
	return s;
}

std::vector<int> parse_cpulist(const std::string &text)
{
	std::vector<int> out;
	const char *p = text.c_str();
	while (*p) {
		while (*p == ',' || *p == ' ' || *p == '\n') ++p;
		if (*p < '0' || *p > '9') break;
		char *e = nullptr;
		long a = strtol(p, &e, 10), b = a;
		if (*e == '-') b = strtol(e + 1, &e, 10);
		for (long c = a; c <= b && c - a < 4096; ++c) out.push_back((int)c);
		p = e;
	}
	return out;
}

int numa_node_count()
{
	const std::vector<int> nodes = parse_cpulist(slurp("/sys/devices/system/node/online"));
	return nodes.empty() ? 1 : (int)nodes.size();
}

int gpu_numa_node(int device)
{
	if (numa_node_count() < 2) return -1;
	char bdf[32] = { 0 };
	if (cudaDeviceGetPCIBusId(bdf, sizeof bdf, device) != cudaSuccess) {
		cudaGetLastError();
		return -1;
	}
	for (char *c = bdf; *c; ++c)
		if (*c >= 'A' && *c <= 'F') *c = (char)(*c - 'A' + 'a');       // sysfs spells the address in lower case
	const std::string s = slurp(std::string("/sys/bus/pci/devices/") + bdf + "/numa_node");
	if (s.empty()) return -1;
	const int n = atoi(s.c_str());
	return n >= 0 ? n : -1;
}

std::vector<int> node_cpus(int node)
{
	std::vector<int> out;
	if (node < 0) return out;
	char path[96];
	snprintf(path, sizeof path, "/sys/devices/system/node/node%d/cpulist", node);
	cpu_set_t allowed;
	CPU_ZERO(&allowed);
	const bool have = sched_getaffinity(0, sizeof allowed, &allowed) == 0;
	for (int c : parse_cpulist(slurp(path)))
		if (c < CPU_SETSIZE && (!have || CPU_ISSET(c, &allowed))) out.push_back(c);
	return out;
}

// <numaif.h> is not in the image either
static constexpr int kMpolDefault = 0, kMpolPreferred = 1;

PreferNode::PreferNode(int node)
{
	if (node < 0 || node >= 1024) return;
	unsigned long mask[16] = { 0 };
	mask[node / (8 * sizeof(unsigned long))] |= 1ul << (node % (8 * sizeof(unsigned long)));
	active = syscall(SYS_set_mempolicy, kMpolPreferred, mask, (unsigned long)(8 * sizeof mask)) == 0;
}
PreferNode::~PreferNode()
{
	if (active) syscall(SYS_set_mempolicy, kMpolDefault, nullptr, 0ul);
}

bool run_on_node(int node)
{
	const std::vector<int> cpus = node_cpus(node);
	if (cpus.empty()) return false;
	cpu_set_t set;
	CPU_ZERO(&set);
	for (int c : cpus) CPU_SET(c, &set);
	return sched_setaffinity(0, sizeof set, &set) == 0;
}

}  // namespace ntsm

// test hooks (not in the public header): the parsing and the no-op behaviour can be checked without a two-socket box
extern "C" int ntsm_numa_parse_cpulist(const char *text, int *out, int cap)
{
	const std::vector<int> v = ntsm::parse_cpulist(text ? text : "");
	for (int i = 0; i < (int)v.size() && i < cap; ++i) out[i] = v[i];
	return (int)v.size();
}
extern "C" int ntsm_numa_nodes(void) { return ntsm::numa_node_count(); }
