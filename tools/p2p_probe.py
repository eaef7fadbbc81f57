"""debug: raw peer-to-peer copy bandwidth between GPU 0 and 1 (one process)"""
import subprocess, torch, time
print(subprocess.run(['nvidia-smi', 'topo', '-m'], capture_output=True, text=True).stdout[:1500])
print('can access peer', torch.cuda.can_device_access_peer(0, 1))
a = torch.empty(1 << 28, dtype=torch.float32, device='cuda:0')   # 1 GiB
b = torch.empty(1 << 28, dtype=torch.float32, device='cuda:1')
for _ in range(3):
    b.copy_(a)
torch.cuda.synchronize(0); torch.cuda.synchronize(1)
t = time.perf_counter()
for _ in range(10):
    b.copy_(a)
torch.cuda.synchronize(0); torch.cuda.synchronize(1)
dt = (time.perf_counter() - t) / 10
print('cudaMemcpyPeer 0->1: %.1f GB/s' % (a.numel() * 4 / dt / 1e9))
