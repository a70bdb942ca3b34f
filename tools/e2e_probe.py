"""GPU probe: where does the end-to-end arm of bench.py lose time vs the resident arm?  (run under gpurun)"""
import sys, time, torch
sys.path.insert(0, ".")
import bench
from pillarnext_b200 import modules, synth, ops
cfg = synth.NUSC
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = modules.build_pillarnext_b(cfg).to(dev).train()
opt = torch.optim.AdamW(model.parameters(), lr=1e-4, betas=(0.9, 0.99), weight_decay=0.01, fused=True)
nb, frames = 4, 6
host = [bench.pin(synth.make_batch([b * frames + f for f in range(frames)], 30000, cfg, kind="lidar", n_boxes=40, sweeps=10)) for b in range(nb)]
resident = [bench.to_device(h, dev) for h in host]
raw = []
for b in range(nb):
    gb, gc = synth.make_gt_batch([b * frames + f for f in range(frames)], 40, cfg)
    raw.append(bench.pin({"points": host[b]["points"], "token": host[b]["token"], "gt_boxes_raw": gb, "gt_classes": gc}))
raw_res = [bench.to_device(h, dev) for h in raw]


def step(ex):
    loss, _ = model(ex)
    loss.backward()
    opt.step()
    opt.zero_grad(set_to_none=True)
    return loss


def timed(fn, n=10):
    for i in range(5):
        fn(i)
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for i in range(n):
        fn(i)
    t1.record()
    torch.cuda.synchronize()
    return t0.elapsed_time(t1) / n


print("resident, precomputed labels   %.2f ms" % timed(lambda i: step(resident[i % nb])))
print("resident, GPU label assignment %.2f ms" % timed(lambda i: step(raw_res[i % nb])))
cs = torch.cuda.Stream(device=dev)


def h2d_step(i, src):
    with torch.cuda.stream(cs):
        ex = bench.to_device(src[i % nb], dev, non_blocking=True)
        ev = torch.cuda.Event(); ev.record(cs)
    torch.cuda.current_stream().wait_event(ev)
    return step(ex)


print("H2D (side stream, not prefetched) + precomputed labels %.2f ms" % timed(lambda i: h2d_step(i, host)))
print("H2D (side stream, not prefetched) + GPU assignment     %.2f ms" % timed(lambda i: h2d_step(i, raw)))
lh = torch.zeros(1).pin_memory()


def item_step(i):
    l = step(raw_res[i % nb])
    lh.copy_(l.detach().reshape(1), non_blocking=True)


print("resident + GPU assignment + async loss copy (no wait)   %.2f ms" % timed(item_step))
evs = [torch.cuda.Event() for _ in range(2)]
lhs = [torch.zeros(1).pin_memory() for _ in range(2)]
state = {"prev": None}


def lazy_item_step(i):
    l = step(raw_res[i % nb])
    s = i & 1
    lhs[s].copy_(l.detach().reshape(1), non_blocking=True)
    evs[s].record()
    if state["prev"] is not None:
        evs[state["prev"]].synchronize()
        float(lhs[state["prev"]][0])
    state["prev"] = s


print("resident + GPU assignment + loss read one step late     %.2f ms" % timed(lazy_item_step))


def bench_loop(n, use_wait_stream=True, side=True):
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    nxt, prev = None, None
    for i in range(n):
        if side:
            if nxt is None:
                with torch.cuda.stream(cs):
                    nxt = (bench.to_device(raw[i % nb], dev, non_blocking=True), torch.cuda.Event())
                    nxt[1].record(cs)
            ex, ev = nxt
            torch.cuda.current_stream().wait_event(ev)
            if i + 1 < n:
                if use_wait_stream:
                    cs.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(cs):
                    nxt = (bench.to_device(raw[(i + 1) % nb], dev, non_blocking=True), torch.cuda.Event())
                    nxt[1].record(cs)
        else:
            ex = bench.to_device(raw[i % nb], dev, non_blocking=True)
        l = step(ex)
        s = i & 1
        lhs[s].copy_(l.detach().reshape(1), non_blocking=True)
        evs[s].record()
        if prev is not None:
            evs[prev].synchronize()
            float(lhs[prev][0])
        prev = s
    evs[prev].synchronize()
    t1.record()
    torch.cuda.synchronize()
    return t0.elapsed_time(t1) / n


for name, kw in (("bench e2e loop (prefetch on side stream + wait_stream)", {}), ("... without wait_stream", {"use_wait_stream": False}),
                 ("... copy on the main stream", {"side": False})):
    bench_loop(3, **kw)
    print("%-60s %.2f ms" % (name, bench_loop(10, **kw)))
    print("%-60s %.2f ms (20 steps)" % (name, bench_loop(20, **kw)))
from pillarnext_b200 import functional as _F; print("pack table rebuilds:", _F._pack_table["rebuilds"], "entries:", len(_F._pack_table["keys"]))
