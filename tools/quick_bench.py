"""Developer timing helper (not the driver's bench): device-resident kernel timings with CUDA events."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import poseidon252_b200 as pb
from poseidon252_b200.scalar import random_limbs_fast

def timeit(fn, stream, iters=5, warm=2):
    for _ in range(warm): fn()
    stream.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    with torch.cuda.stream(stream):
        for a, b in evs:
            a.record(stream); fn(); b.record(stream)
    stream.synchronize()
    return sorted(a.elapsed_time(b) for a, b in evs)

def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else (1 << 20)
    st = torch.cuda.Stream()
    eng = pb.Engine(0, stream=st.cuda_stream)
    rng = np.random.default_rng(0)
    with torch.cuda.stream(st):
        leaves = torch.from_numpy(random_limbs_fast(rng, (n, 4)).view(np.int64)).cuda()
        out = torch.empty((n, 1, 4), dtype=torch.int64, device="cuda")
        states = torch.from_numpy(random_limbs_fast(rng, (n, 5)).view(np.int64)).cuda()
    st.synchronize()
    t = timeit(lambda: pb.Hash.digest_batch(pb.Domain.Merkle4, leaves, engine=eng, out=out, async_=True), st)
    print("merkle4 digest n=%d: median %.3f ms  -> %.3e perm/s" % (n, t[len(t)//2], n / (t[len(t)//2] * 1e-3)))
    t = timeit(lambda: eng.permute_batch_inplace(states, async_=True), st)
    print("permute n=%d: median %.3f ms  -> %.3e perm/s" % (n, t[len(t)//2], n / (t[len(t)//2] * 1e-3)))
    nd = n // 16
    t = timeit(lambda: eng.permute_batch(states[:nd], dense=True, async_=True), st, iters=3, warm=1)
    print("dense permute n=%d: median %.3f ms  -> %.3e perm/s" % (nd, t[len(t)//2], nd / (t[len(t)//2] * 1e-3)))
    msg = torch.from_numpy(random_limbs_fast(rng, (n, 2)).view(np.int64)).cuda()
    sec = torch.from_numpy(random_limbs_fast(rng, (n, 2)).view(np.int64)).cuda()
    non = torch.from_numpy(random_limbs_fast(rng, (n,)).view(np.int64)).cuda()
    cip = torch.empty((n, 3, 4), dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    t = timeit(lambda: pb.encrypt_batch(msg, sec, non, engine=eng, out=cip, async_=True), st)
    print("encrypt L=2 n=%d: median %.3f ms  -> %.3e perm/s" % (n, t[len(t)//2], 2 * n / (t[len(t)//2] * 1e-3)))
    # host path e2e
    h_in = random_limbs_fast(rng, (n, 4)); 
    t0 = time.time(); pb.Hash.digest_batch(pb.Domain.Merkle4, h_in, engine=eng); t1 = time.time()
    t0 = time.time(); pb.Hash.digest_batch(pb.Domain.Merkle4, h_in, engine=eng); t1 = time.time()
    print("host e2e (pageable) n=%d: %.3f ms -> %.3e /s" % (n, (t1 - t0) * 1e3, n / (t1 - t0)))
    print(os.popen("nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw --format=csv,noheader").read())

if __name__ == "__main__":
    main()
