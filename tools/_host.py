import sys, time, cProfile, pstats, torch
sys.path.insert(0, '/root/repo')
from mustafar_b200.attention import MustafarKVCache
b,hkv,g,T,s=1,32,1,4096,0.5
caches=[]
for l in range(8):
    k = torch.randn(b, hkv, T, 128, device='cuda', dtype=torch.float16); v = torch.randn_like(k)
    c = MustafarKVCache(b, hkv, g, T+2048, s, s); c.prefill(k, v); c.win_len -= 1; caches.append(c)
q = torch.randn(b, hkv, 1, 128, device='cuda', dtype=torch.float16); kn = torch.randn_like(q); vn = torch.randn_like(q); out = torch.empty_like(q)
for c in caches: c.decode_step(q, kn, vn, out=out)
torch.cuda.synchronize()
t0=time.perf_counter()
n=0
for it in range(25):
    for c in caches:
        c.decode_step(q, kn, vn, out=out); n+=1
t1=time.perf_counter()
torch.cuda.synchronize(); t2=time.perf_counter()
print(f'host time per decode_step (queue not full): {(t1-t0)/n*1e6:.2f} us; incl. drain {(t2-t0)/n*1e6:.2f} us')
pr=cProfile.Profile(); pr.enable()
for it in range(25):
    for c in caches: c.decode_step(q, kn, vn, out=out)
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats('tottime').print_stats(12)
