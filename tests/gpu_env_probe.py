"""First-gpurun checklist: device facts written to gpurun_out/env_probe.json."""
import json, os, subprocess, sys
import torch
p = torch.cuda.get_device_properties(0)
info = dict(name=p.name, sms=p.multi_processor_count, mem_gb=p.total_memory / 2**30, l2_mb=p.L2_cache_size / 2**20,
            max_threads_per_sm=p.max_threads_per_multi_processor, arch_list=torch.cuda.get_arch_list(),
            cpus=os.cpu_count(), torch_threads=torch.get_num_threads(), n_gpus=torch.cuda.device_count())
try:
    info["nvidia_smi"] = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm,power.draw,power.limit", "--format=csv,noheader"],
                                        capture_output=True, text=True).stdout.strip()
except Exception as e:
    info["nvidia_smi"] = repr(e)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(info, open("gpurun_out/env_probe.json", "w"), indent=1)
print(info)
