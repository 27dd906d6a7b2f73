"""Convert between a TensorFlow-1 checkpoint of the reference (`tf.train.Saver`, agents/models.py:83-108) and the
name-keyed `checkpoint-<step>.npz` this repo reads / writes (deeprl_signal_control_b200/agents/checkpoint.py).

Runs on a host that HAS TensorFlow (1.x, or 2.x through tf.compat.v1); this container has none, so the script is not
exercised by the test suite — the tensor NAMES and SHAPES it relies on are (tests/test_learner_reference_golden_cpu.py).

  python scripts/convert_tf_checkpoint.py to-npz   <model_dir>/checkpoint-1000080  out_dir/checkpoint-1000080.npz
  python scripts/convert_tf_checkpoint.py to-tf    in_dir/checkpoint-720.npz       <model_dir>/checkpoint-720
"""
import sys

import numpy as np


def to_npz(ckpt_prefix, out_npz):
    import tensorflow as tf
    reader = tf.train.load_checkpoint(ckpt_prefix)
    names = [n for n in reader.get_variable_to_shape_map() if "RMSProp" not in n and "Adam" not in n and "power" not in n]
    np.savez(out_npz, **{n: reader.get_tensor(n) for n in sorted(names)})
    print("wrote %d tensors to %s" % (len(names), out_npz))


def to_tf(in_npz, ckpt_prefix):
    import tensorflow as tf
    tf1 = tf.compat.v1 if hasattr(tf, "compat") else tf
    tf1.disable_eager_execution()
    z = np.load(in_npz)
    names = [k for k in z.files if not k.startswith("__b200__/")]
    with tf1.Graph().as_default():
        vs = [tf1.get_variable(n, initializer=tf1.constant(z[n])) for n in names]
        with tf1.Session() as sess:
            sess.run(tf1.global_variables_initializer())
            prefix, step = ckpt_prefix.rsplit("-", 1)
            tf1.train.Saver(vs).save(sess, prefix, global_step=int(step))
    print("wrote %d variables to %s" % (len(names), ckpt_prefix))


if __name__ == "__main__":
    if len(sys.argv) != 4 or sys.argv[1] not in ("to-npz", "to-tf"):
        sys.exit(__doc__)
    (to_npz if sys.argv[1] == "to-npz" else to_tf)(sys.argv[2], sys.argv[3])
