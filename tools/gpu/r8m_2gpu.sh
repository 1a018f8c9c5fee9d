set -x
nvidia-smi topo -m 2>&1 | head -8
python - <<'PY'
import os, pynvml, torch
pynvml.nvmlInit()
print("allowed cpus", len(os.sched_getaffinity(0)), sorted(os.sched_getaffinity(0))[:4], "...", os.cpu_count())
for i in range(torch.cuda.device_count()):
    pr = torch.cuda.get_device_properties(i)
    bus = "%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
    h = pynvml.nvmlDeviceGetHandleByPciBusId(bus)
    m = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count()+63)//64)
    cpus = {64*k+b for k,w in enumerate(m) for b in range(64) if (int(w)>>b)&1}
    print(i, bus, len(cpus), len(cpus & os.sched_getaffinity(0)))
PY
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-parity ) > gpurun_out/r8m_bench2.json 2> gpurun_out/r8m_bench2.err; tail -c 1800 gpurun_out/r8m_bench2.json; grep -v "^$" gpurun_out/r8m_bench2.err | tail -6
