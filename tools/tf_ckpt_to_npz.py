#!/usr/bin/env python
"""Convert between TensorFlow checkpoint V2 bundles and the `.npz` exchange file (TF-Slim variable names), both ways,
WITHOUT TensorFlow (uses unsuperviseddeephomographyral2018_b200/tf_checkpoint.py):

  python tools/tf_ckpt_to_npz.py to-npz  <ckpt prefix, e.g. models/synthetic_models/l1_loss_normalize/model.ckpt-150000>  out.npz
  python tools/tf_ckpt_to_npz.py to-ckpt in.npz  <ckpt prefix>

With TensorFlow 1.x at hand the same npz is `{v.name[:-2]: sess.run(v) for v in tf.global_variables()}`."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unsuperviseddeephomographyral2018_b200 import tf_checkpoint as tfc


def main(argv):
    if len(argv) != 4 or argv[1] not in ("to-npz", "to-ckpt"):
        print(__doc__); return 2
    if argv[1] == "to-npz":
        np.savez(argv[3], **tfc.read_checkpoint(argv[2]))
    else:
        with np.load(argv[2]) as z:
            tfc.write_checkpoint(argv[3], {k: z[k] for k in z.files})
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
