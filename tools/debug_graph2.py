import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

def probe(ng, tile_n, use_graph):
    import torch
    from pql_b200 import _kernels as K, _lib
    dev = "cuda:0"
    _lib.check(_lib.load().pqlb_init(), "init")
    M, N, Kd = 512, 512, 104
    g = torch.Generator(device=dev).manual_seed(0)
    groups, outs, refs, keep = [], [], [], []
    for i in range(ng):
        a = torch.randn(M, Kd, device=dev, generator=g); w = torch.randn(N, Kd, device=dev, generator=g) * 0.1
        b = torch.randn(N, device=dev, generator=g); o = torch.zeros(M, N, device=dev)
        groups.append(dict(a=K.addr(a), lda=Kd, b=K.addr(w), ldb=Kd, bias=K.addr(b), out=K.addr(o), ldo=N))
        outs.append(o); refs.append(a @ w.t() + b); keep += [a, w, b]
    call = K.Gemm(M, N, Kd, groups, epilogue=K.EPI_BIAS, tile_n=tile_n)
    torch.cuda.synchronize()
    if use_graph:
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            call()
        gr.replay()
    else:
        call()
    torch.cuda.synchronize()
    err = max((o - r).abs().max().item() for o, r in zip(outs, refs))
    print("OK", ng, tile_n, use_graph, "maxerr", err, flush=True)

if __name__ == "__main__":
    if len(sys.argv) > 1:
        probe(int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]))
    else:
        for ng in (1, 2, 3, 4):
            for tile in (64, 256):
                for gph in (0, 1):
                    r = subprocess.run([sys.executable, __file__, str(ng), str(tile), str(gph)], capture_output=True, text=True)
                    print(ng, tile, gph, r.returncode, (r.stdout.strip().splitlines() or [r.stderr.strip()[-300:]])[-1], flush=True)
