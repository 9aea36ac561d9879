# 8-GPU box: concurrent PCIe ceiling (plain pinned copies, N ranks at once) and the bench at N = 8, 4.
cd /root/repo
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi topo -m > gpurun_out/r2_topo.txt 2>&1
(lscpu | head -25; numactl -H 2>/dev/null; for d in /sys/bus/pci/devices/*; do c=$(cat $d/class 2>/dev/null); case $c in 0x0302*|0x0300*) echo "$d numa=$(cat $d/numa_node)";; esac; done) > gpurun_out/r2_host.txt 2>&1
$TR --nproc-per-node 8 --master-port 29608 tools/pcie_concurrent.py --seconds 1 > gpurun_out/r2_pcie_n8.json 2> gpurun_out/r2_pcie_n8.err
$TR --nproc-per-node 8 --master-port 29618 tools/pcie_concurrent.py --bind --seconds 1 > gpurun_out/r2_pcie_bind_n8.json 2> gpurun_out/r2_pcie_bind_n8.err
$TR --nproc-per-node 4 --master-port 29614 tools/pcie_concurrent.py --bind --seconds 1 > gpurun_out/r2_pcie_bind_n4.json 2> gpurun_out/r2_pcie_bind_n4.err
$TR --nproc-per-node 2 --master-port 29612 tools/pcie_concurrent.py --bind --seconds 1 > gpurun_out/r2_pcie_bind_n2.json 2> gpurun_out/r2_pcie_bind_n2.err
python tools/pcie_concurrent.py --seconds 1 > gpurun_out/r2_pcie_n1.json 2> gpurun_out/r2_pcie_n1.err
$TR --nproc-per-node 8 --master-port 29808 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2_bench_8gpu.json 2> gpurun_out/r2_bench_8gpu.err
$TR --nproc-per-node 4 --master-port 29804 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/r2_bench_4gpu.json 2> gpurun_out/r2_bench_4gpu.err
tail -c 300 gpurun_out/r2_bench_8gpu.err
cat gpurun_out/r2_pcie_*.json
