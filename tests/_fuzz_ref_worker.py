"""Reference side of the GPU fuzz (CPU only, never touches CUDA): `python _fuzz_ref_worker.py kind seed cases out.pkl`.
Every case runs in a forked child -- the reference corrupts its heap / loops on some legal parameter sets -- with a
time limit, against the AddressSanitizer build of the reference when the caller preloads libasan (AFB200_FUZZ_ASAN=1);
the pickle holds [(params, result dict | 'crash')]."""
import os
import pickle
import signal
import sys

HERE = os.path.dirname(os.path.realpath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

import numpy as np  # noqa: E402

from _fuzz_cases import compute, gen  # noqa: E402
from oracle import ref_lib as R  # noqa: E402


def in_child(fn, seconds=30):
    r, w = os.pipe()
    pid = os.fork()
    if pid == 0:
        os.close(r)
        try:
            import faulthandler
            faulthandler.disable()
            signal.alarm(seconds)
            devnull = os.open(os.devnull, os.O_WRONLY)
            os.dup2(devnull, 1)
            os.dup2(devnull, 2)
            os.write(w, pickle.dumps(fn()))
        finally:
            os._exit(0)
    os.close(w)
    data = b""
    while True:
        c = os.read(r, 1 << 20)
        if not c:
            break
        data += c
    os.close(r)
    os.waitpid(pid, 0)
    return pickle.loads(data) if data else "crash"


def main():
    kind, seed, cases, out = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
    # under LD_PRELOAD=libasan the AddressSanitizer build: a case over which the reference overruns a buffer aborts its child
    # and is reported as 'crash' instead of a result computed over a corrupted heap
    asan = os.environ.get("AFB200_FUZZ_ASAN") == "1" and os.path.exists(R.REF_ASAN_PATH)
    lib = R.get_ref_lib(asan=asan)
    rng = np.random.default_rng(seed)
    res = []
    for _ in range(cases):
        a = gen(kind, rng)
        res.append((a, in_child(lambda: compute(kind, a, lib))))
    with open(out, "wb") as f:
        pickle.dump(res, f)


if __name__ == "__main__":
    main()
