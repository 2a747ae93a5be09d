# development aid: per-phase clock64 timing of the QR kernel (built with -DSB200_QR_TIMING)
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import strumpack_b200 as sb
variant = sys.argv[2] if len(sys.argv) > 2 else ""
sb._SO = os.path.join(ROOT, "scripts", "libsb200_timing%s.so" % (("_" + variant) if variant else ""))
import numpy as np
n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
pts = np.random.default_rng(42).random((2, n))
o = sb.default_options(type=sb.SP_TYPE_HSS, rel_tol=1e-4, abs_tol=1e-10, leaf_size=256)
H, perm, p = sb.HSSMatrix.from_kernel(pts, sb.KERNEL_GAUSS, 0.1, 1.0, o)
H.factor()
