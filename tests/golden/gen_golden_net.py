"""Net fixtures (build container only): converts /root/reference/good_model.h5 (Keras save_weights) into
tests/golden/good_model_weights.npz with the product's own HDF5 reader, and records the float64
restatement's outputs (oracle/net_ref.py) on positions from env_golden.npz.  PARITY UNPINNED for the net:
the reference's arithmetic is Keras/TF (un-vendored); see oracle/net_ref.py."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import net_ref  # noqa: E402
from chinesecheckersagent_b200 import h5lite  # noqa: E402

REF = os.environ.get("CCX_REFERENCE_DIR", "/root/reference")


def main():
    w = h5lite.read_weights(os.path.join(REF, "good_model.h5"))
    np.savez_compressed(os.path.join(HERE, "good_model_weights.npz"), **w)
    env = np.load(os.path.join(HERE, "env_golden.npz"))
    planes = env["planes"][::15][:256]                      # 256 positions: random, randomised, greedy, terminal
    logits, v = net_ref.forward(w, planes, np.float64)
    p = net_ref.softmax64(logits)
    # start position known answers for all three shipped weight files
    ka = {}
    for f in ("good_model.h5", "good_model2.h5", "version0016-weights.h5"):
        ww = h5lite.read_weights(os.path.join(REF, f))
        pp, vv = net_ref.predict(ww, env["planes"][:1], np.float32)
        ka[f] = (float(vv[0]), [int(i) for i in np.argsort(-pp[0])[:5]])
        print(f, ka[f])
    np.savez_compressed(os.path.join(HERE, "net_golden.npz"), planes=planes, logits=logits, v=v)
    print("net_golden: %d positions; max |logit| %.3f" % (len(planes), np.abs(logits).max()))


if __name__ == "__main__":
    main()
