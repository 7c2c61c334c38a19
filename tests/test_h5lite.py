"""§8f f4: the HDF5 subset reader/writer (Keras save_weights files, data-for-iter files).  CPU only."""
import os

import numpy as np
import pytest

from conftest import GOLDEN

REF_H5 = "/root/reference/good_model.h5"


def test_weights_round_trip(tmp_path):
    from chinesecheckersagent_b200 import h5lite
    w = {k: np.asarray(v, dtype=np.float32) for k, v in np.load(os.path.join(GOLDEN, "good_model_weights.npz")).items()}
    path = h5lite.write_weights(str(tmp_path / "version0001.h5"), w)
    back = h5lite.read_weights(path)
    assert set(back) == set(w) and all(np.array_equal(back[k], w[k]) for k in w)
    f = h5lite._File(open(path, "rb").read())
    root = f.attributes(f.root_header)
    assert [n.decode() for n in root["layer_names"]] == h5lite.keras_layer_order()
    assert root["keras_version"] == b"2.1.6" and root["backend"] == b"tensorflow"
    ent = f.group_entries(f.root_header)
    assert [n.decode() for n in f.attributes(ent["batch_normalization_7"])["weight_names"]] == [
        "batch_normalization_7/gamma:0", "batch_normalization_7/beta:0", "batch_normalization_7/moving_mean:0",
        "batch_normalization_7/moving_variance:0"]
    assert len(f.attributes(ent["add_3"])["weight_names"]) == 0


def test_train_data_round_trip(tmp_path):
    from chinesecheckersagent_b200 import utils
    rng = np.random.default_rng(0)
    bx = rng.integers(0, 7, (37, 7, 7, 7)).astype(np.float64)
    pi = rng.random((37, 294))
    v = rng.integers(-1, 2, 37)
    path = utils.save_train_data(bx, pi, v, 12, directory=str(tmp_path))
    assert path.endswith("data-for-iter-12.h5")
    b2, p2, v2 = utils.load_train_data(path)
    assert np.array_equal(b2, bx) and np.array_equal(p2, pi) and np.array_equal(v2, v)


@pytest.mark.skipif(not os.path.exists(REF_H5), reason="reference checkout only exists in the build container")
def test_written_file_mirrors_a_real_keras_file(tmp_path):
    """same groups, same layer order, same weight_names, same tensors as the file Keras 2.1.6 wrote for the reference"""
    from chinesecheckersagent_b200 import h5lite
    real = h5lite._File(open(REF_H5, "rb").read())
    w = h5lite.read_weights(REF_H5)
    mine = h5lite._File(open(h5lite.write_weights(str(tmp_path / "w.h5"), w), "rb").read())
    assert [n for n in real.attributes(real.root_header)["layer_names"]] == [n for n in mine.attributes(mine.root_header)["layer_names"]]
    er, em = real.group_entries(real.root_header), mine.group_entries(mine.root_header)
    assert set(er) == set(em)
    for name in er:
        assert list(real.attributes(er[name])["weight_names"]) == list(mine.attributes(em[name])["weight_names"]), name
    tr, tm = h5lite.read_tree(REF_H5), h5lite.read_tree(str(tmp_path / "w.h5"))
    assert set(tr) == set(tm) and all(np.array_equal(tr[k], tm[k]) and tr[k].dtype == tm[k].dtype for k in tr)


def test_weights_of_a_second_model_in_the_process_load_by_position(tmp_path):
    """Keras numbers auto-named layers per process: a weights file saved from the second ResidualCNN built in a process holds
    conv2d_31..60 / batch_normalization_31..60 / dense_2, and a full-model save keeps everything under /model_weights.  Keras'
    load_weights matches by position, so the reference loads both; read_weights must too."""
    import re

    from chinesecheckersagent_b200 import h5lite
    w = {k: np.asarray(v, dtype=np.float32) for k, v in np.load(os.path.join(GOLDEN, "good_model_weights.npz")).items()}

    def shifted(name):
        m = re.fullmatch(r"(conv2d|batch_normalization)_(\d+)", name)
        if m:
            return "%s_%d" % (m.group(1), int(m.group(2)) + 30)
        return "dense_2" if name == "dense_1" else name
    tree = {}
    for k, v in w.items():
        layer, param = k.split("/")
        tree["model_weights/%s/%s/%s:0" % (shifted(layer), shifted(layer), param)] = v
    path = h5lite.write_tree(str(tmp_path / "second_model.h5"), tree)
    back = h5lite.read_weights(path)
    assert set(back) == set(w) and all(np.array_equal(back[k], w[k]) for k in w)
