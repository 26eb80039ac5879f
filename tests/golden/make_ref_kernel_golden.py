"""Generates tests/golden/ref_kernels_golden.npz: the outputs of the REFERENCE's own kernels.cu (compiled
verbatim to oracle/_ref/helios_ref.cubin, launched with the block/grid shapes of source/computation.py) at
every launch site of the hot path, on the seeded tiny cases of tests/golden/cases.py.

Needs a GPU and the prebuilt cubin (which needs /root/reference at build time), so it runs on the B200 box:

    gpurun -- 'python tests/golden/make_ref_kernel_golden.py gpurun_out/ref_kernels_golden.npz'

and the result is committed as tests/golden/ref_kernels_golden.npz.  The CPU suite (tests/test_oracle_golden.py)
pins the NumPy oracle against these vectors; the GPU suite compares the product kernels with the same
reference kernels live.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def main(out_path):
    import cases
    from helios_b200 import runtime, synthetic
    from oracle import ref_gpu
    ctx = runtime.default_context()
    ref = ref_gpu.RefCompute(ctx.device)
    out = {}
    for case in cases.CASES:
        q = cases.build(case, ctx)
        synthetic.upload(q)
        ctx.synchronize()
        counter = [0]

        def after(label, names, case=case, q=q):
            ctx.synchronize()
            ref.mod.synchronize()
            for n in names:
                dev = getattr(q, "dev_" + n, None)
                if dev is None:
                    continue
                out["%s|%03d|%s|%s" % (case, counter[0], label, n)] = dev.get().copy()
            counter[0] += 1

        cases.run_case(case, q, ref, cases.DeviceIO, after, ref_signature=True)
        print("%-20s %3d launch sites recorded" % (case, counter[0]))
    np.savez_compressed(out_path, **out)
    print("wrote %s: %d arrays, %.1f kB" % (out_path, len(out), os.path.getsize(out_path) / 1e3))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_kernels_golden.npz"))
