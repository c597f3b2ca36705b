import torch, time
d = torch.empty(1_490_000_000, dtype=torch.uint8, device="cuda")
h = torch.empty(1_490_000_000, dtype=torch.uint8).pin_memory()
h2 = torch.empty(720_000_000, dtype=torch.uint8).pin_memory(); d2 = torch.empty(720_000_000, dtype=torch.uint8, device="cuda")
for _ in range(2): h.copy_(d, non_blocking=True); torch.cuda.synchronize()
t=time.perf_counter(); 
for _ in range(3): h.copy_(d, non_blocking=True)
torch.cuda.synchronize(); dt=(time.perf_counter()-t)/3; print("D2H %.1f GB/s" % (1.49/dt))
t=time.perf_counter()
for _ in range(3): d2.copy_(h2, non_blocking=True)
torch.cuda.synchronize(); dt=(time.perf_counter()-t)/3; print("H2D %.1f GB/s" % (0.72/dt))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
t=time.perf_counter()
for _ in range(3):
    with torch.cuda.stream(s1): h.copy_(d, non_blocking=True)
    with torch.cuda.stream(s2): d2.copy_(h2, non_blocking=True)
torch.cuda.synchronize(); dt=(time.perf_counter()-t)/3; print("both: D2H %.1f GB/s + H2D %.1f GB/s" % (1.49/dt, 0.72/dt))
import subprocess; print(subprocess.run(["nvidia-smi","--query-gpu=pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max","--format=csv"],capture_output=True,text=True).stdout)
