#!/bin/bash
# round 2, call 1: platform topology, link ceiling at N=1 and N=2, e2e re-baseline at N=2
mkdir -p gpurun_out
{
  echo "== nproc"; nproc; echo "== lscpu"; lscpu | head -40
  echo "== numactl"; numactl -H 2>&1 | head -20
  echo "== nvidia-smi topo"; nvidia-smi topo -m 2>&1
  echo "== nvidia-smi -L"; nvidia-smi -L
  echo "== pci numa"; for d in /sys/bus/pci/devices/*; do c=$(cat $d/class 2>/dev/null); if [[ "$c" == 0x0302* || "$c" == 0x0300* ]]; then echo "$d $(cat $d/numa_node) $(cat $d/local_cpulist 2>/dev/null)"; fi; done
  echo "== cgroup cpu"; cat /sys/fs/cgroup/cpu.max 2>/dev/null; cat /sys/fs/cgroup/cpuset.cpus.effective 2>/dev/null
  echo "== nodes"; ls /sys/devices/system/node/ 2>/dev/null
  echo "== mem"; free -g | head -2
  echo "== pcie link"; nvidia-smi --query-gpu=index,pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max --format=csv
} > gpurun_out/topo.txt 2>&1
python tools/link_probe.py > gpurun_out/probe_n1.json 2> gpurun_out/probe_n1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/link_probe.py > gpurun_out/probe_n2.json 2> gpurun_out/probe_n2.err
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_n1_base.json 2> gpurun_out/bench_n1_base.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_n2_base.json 2> gpurun_out/bench_n2_base.err
echo done
