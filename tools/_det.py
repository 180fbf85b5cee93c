import sys, torch
sys.path.insert(0, '/root/repo')
from mustafar_b200.attention import MustafarKVCache
b,hkv,g,T,s=4,8,4,8192,0.7
torch.manual_seed(0)
k = torch.randn(b, hkv, T, 128, device='cuda', dtype=torch.float16)
v = torch.randn(b, hkv, T, 128, device='cuda', dtype=torch.float16)
q = torch.randn(b, hkv*g, 1, 128, device='cuda', dtype=torch.float16)
c = MustafarKVCache(b, hkv, g, T, s, s, pdl=False)
c.prefill(k, v)
for it in range(3):
    o = c.attend(q); torch.cuda.synchronize()
    p = c._p; units=b*hkv; ns=p.n_split
    cb=(units*4+255)//256*256
    parts = c._ws[cb:cb+units*ns*g*132*4].view(torch.float32).view(units, ns, g, 132)
    nan_o = torch.isnan(parts[...,:128]).any(-1); nan_m = torch.isnan(parts[...,128]); nan_l=torch.isnan(parts[...,129])
    print('call',it,'out nan heads',int(torch.isnan(o).any(-1).sum()),'partials nan o:',nan_o.nonzero()[:5].tolist(),'m:',nan_m.nonzero()[:5].tolist(),'l:',nan_l.nonzero()[:5].tolist(), 'inf m', torch.isinf(parts[...,128]).nonzero()[:4].tolist(), 'ncsplit', ns - (c.win_len+63)//64)
    if nan_o.any():
        u,sp,gg = nan_o.nonzero()[0].tolist()
        row = parts[u,sp,gg,:128]
        print('  nan channels', torch.isnan(row).nonzero().flatten()[:20].tolist(), 'm,l', parts[u,sp,gg,128].item(), parts[u,sp,gg,129].item())
