import sys, os, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1:
    import torch
    from coma_b200.inpaint import nn
    S, L, d, heads = map(int, sys.argv[1:5])
    dev = torch.device("cuda:0"); B = 2; C = d * heads
    g = torch.Generator(device=dev).manual_seed(0)
    q = torch.randn((B * S, C), device=dev, generator=g).half(); k = torch.randn((B * L, C), device=dev, generator=g).half(); v = torch.randn((B * L, C), device=dev, generator=g).half()
    Lp = nn.rup(L)
    vt = torch.empty((B, heads, d, Lp), dtype=torch.float16, device=dev); o = torch.empty((B * S, C), dtype=torch.float16, device=dev)
    nn.call("coma_transpose_heads_f16", v.data_ptr(), B, L, heads, d, C, vt.data_ptr(), Lp, nn._stream())
    nn.call("coma_attention_fwd_f16", q.data_ptr(), k.data_ptr(), vt.data_ptr(), B, heads, S, L, d, C, C, Lp, float(d ** -0.5), o.data_ptr(), C, nn._stream())
    torch.cuda.synchronize()
    sp = lambda t, n: t.float().view(B, n, heads, d).permute(0, 2, 1, 3)
    ref = torch.softmax(sp(q, S) @ sp(k, L).transpose(-1, -2) * d ** -0.5, -1) @ sp(v, L)
    ref = ref.permute(0, 2, 1, 3).reshape(B * S, C)
    print("ok", sys.argv[1:5], "max err", (o.float() - ref).abs().max().item())
else:
    for case in [(16, 16, 16, 2), (64, 64, 8, 2), (256, 256, 96, 2), (128, 200, 176, 1), (64, 64, 56, 2)]:
        try:
            r = subprocess.run([sys.executable, __file__] + [str(c) for c in case], capture_output=True, text=True, timeout=40)
            print(r.stdout.strip() or ("FAILED " + str(case) + " " + r.stderr.strip()[-300:]), flush=True)
        except subprocess.TimeoutExpired:
            print("HANG", case, flush=True)
