#!/bin/bash
# Host-side facts that bound the end-to-end (host memory) path: CPUs, NUMA nodes, GPU/PCIe topology.
echo "== nproc: $(nproc)"; lscpu | egrep 'Model name|Socket|NUMA|Thread|Core|Hypervisor' 
echo "== NUMA nodes:"; ls -d /sys/devices/system/node/node* 2>/dev/null; for n in /sys/devices/system/node/node*; do echo "$n cpus=$(cat $n/cpulist) $(grep MemTotal $n/meminfo)"; done
echo "== nvidia-smi topo -m"; nvidia-smi topo -m
echo "== PCIe link per GPU"; nvidia-smi --query-gpu=index,pci.bus_id,pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max --format=csv
for d in /sys/bus/pci/devices/*; do if [ -f $d/numa_node ] && grep -q 0x10de $d/vendor 2>/dev/null; then echo "$(basename $d) numa_node=$(cat $d/numa_node)"; fi; done | head -20
grep -i iommu /proc/cmdline; dmesg 2>/dev/null | grep -i -m3 iommu
free -g | head -2
