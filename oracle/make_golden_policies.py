"""Golden vectors for the reference's other two pruning policies, produced by the REFERENCE'S OWN functions
(run in the build container only; /root/reference does not exist on the GPU box).

  * output-aware key pruning:  `dh_prune_key` of /root/reference/models/llama_mustafar_Kt_Opa_Vt_Mag.py:65-178
  * channel-wise value pruning: `dh_prune_value` of /root/reference/models/llama_mustafar_Kt_Mag_Vc_Mag.py:107-170

Both model files fail to import under transformers 5.x, so the two methods are extracted with `ast` and executed verbatim
against a stub `self` that carries the attributes they read; nothing is copied into this repository.
Writes tests/golden/policy_opa_*.npz and tests/golden/policy_vc_*.npz.     Usage: python oracle/make_golden_policies.py
"""
import ast
import os
import sys
import types
from typing import Optional  # noqa: F401  (the extracted signatures use it)

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def extract_method(path, name):
    tree = ast.parse(open(os.path.join(REF, path)).read())
    fn = next(n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef) and n.name == name)
    ns = {"torch": torch, "Optional": Optional, "DEBUG": False}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), f"<ref {path}:{name}>", "exec"), ns)
    return ns[name]


def cut_is_tied(score, n_keep):
    """True for rows whose n_keep-th and (n_keep+1)-th highest scores are equal (the reference's choice is then arbitrary)."""
    s = np.sort(score.astype(np.float32), axis=-1)[..., ::-1]
    return s[..., n_keep - 1] == s[..., n_keep]


def main():
    os.makedirs(OUT, exist_ok=True)
    opa = extract_method("models/llama_mustafar_Kt_Opa_Vt_Mag.py", "dh_prune_key")
    vc = extract_method("models/llama_mustafar_Kt_Mag_Vc_Mag.py", "dh_prune_value")

    # ---- output-aware keys: prefill call (iteration 0) followed by decode calls on a sliding window ----
    for name, seed, b, hkv, groups, t, s, gs in [("mha_s50", 21, 2, 2, 1, 96, 0.5, 32), ("gqa4_s70", 22, 1, 2, 4, 64, 0.7, 16)]:
        g = torch.Generator().manual_seed(seed)
        me = types.SimpleNamespace(k_sparsity=s, group_size=gs, num_heads=hkv * groups, num_key_value_groups=groups,
                                   calculate_sparsity=lambda x: 0.0)
        key = torch.randn(b, hkv, t, 128, generator=g).half()
        q = torch.randn(b, hkv * groups, t, 128, generator=g).half()
        acc = torch.zeros(b, hkv, gs + 1, 128, dtype=torch.float16)
        _, pruned = opa(me, 0, key, q, acc)
        n_keep = int(128 * (1 - s))
        w = torch.mean(torch.abs(q[:, :, -gs:, :]), dim=-2).view(b, hkv, groups, 128).sum(dim=-2)  # `:98-100`
        score = torch.abs(w[:, :, None, :] * key)
        out = {"key": key.numpy(), "q": q.numpy(), "w": w.numpy(), "pruned": pruned.numpy(), "acc_after_prefill": acc.numpy(),
               "tied_rows": cut_is_tied(score.numpy(), n_keep), "sparsity": np.float64(s), "group_size": np.int64(gs),
               "groups": np.int64(groups)}
        # decode: the window holds the gs + 1 newest keys; every step scores the gs newest, prunes the oldest
        steps = 6
        win = key[:, :, -(gs + 1):, :].clone()
        d_win, d_q, d_acc, d_out, d_tied = [], [], [], [], []
        for _ in range(steps):
            k_new = torch.randn(b, hkv, 1, 128, generator=g).half()
            win = torch.cat([win[:, :, 1:, :], k_new], dim=2)
            q1 = torch.randn(b, hkv * groups, 1, 128, generator=g).half()
            d_win.append(win.numpy().copy())
            d_q.append(q1.numpy().copy())
            d_acc.append(acc.numpy().copy())
            # what the call sorts: the oldest row's accumulated score / group_size (`:131-139`; the += touches newer rows only)
            d_tied.append(cut_is_tied((acc[:, :, 0:1, :] / gs).numpy(), n_keep))
            _, row = opa(me, 1, win, q1, acc)
            d_out.append(row.numpy().copy())
        out.update(dec_window=np.stack(d_win), dec_q=np.stack(d_q), dec_acc_before=np.stack(d_acc), dec_pruned=np.stack(d_out),
                   dec_tied=np.stack(d_tied), acc_final=acc.numpy())
        path = os.path.join(OUT, f"policy_opa_{name}.npz")
        np.savez_compressed(path, **out)
        print("wrote", path, "tied prefill rows:", int(out["tied_rows"].sum()), "tied decode rows:", int(out["dec_tied"].sum()))

    # ---- channel-wise values ----
    for name, seed, shape, s, gs, kind in [("randn_s50_g32", 31, (2, 2, 128, 128), 0.5, 32, "randn"),
                                          ("randn_s70_g32", 32, (1, 3, 64, 128), 0.7, 32, "randn"),
                                          ("ties_s50_g64", 33, (1, 2, 128, 128), 0.5, 64, "ties"),
                                          ("randn_s90_g128", 34, (1, 1, 256, 128), 0.9, 128, "randn"),
                                          ("randn_s001_g32", 35, (1, 1, 32, 128), 0.001, 32, "randn")]:
        g = torch.Generator().manual_seed(seed)
        x = torch.randn(*shape, generator=g)
        if kind == "ties":
            x = torch.round(x * 4) / 4
        x = x.half()
        me = types.SimpleNamespace(v_sparsity=s, group_size=gs, calculate_sparsity=lambda x: 0.0)
        y = vc(me, x)
        path = os.path.join(OUT, f"policy_vc_{name}.npz")
        np.savez_compressed(path, x=x.numpy(), y=y.numpy(), sparsity=np.float64(s), group_size=np.int64(gs))
        print("wrote", path)


if __name__ == "__main__":
    sys.exit(main())
