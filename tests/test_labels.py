"""Label codec: the reference's known answers (util/test_dataProcessor.py:132-229)
and the golden outputs of the reference's own codec."""
import numpy as np
import pytest

from conftest import golden


def test_known_answers_from_reference_tests(pkg):
    cm = pkg.ENGLISH_CHAR_MAP
    assert len(cm) == 80 and cm[79] == "_"
    assert pkg.get_str_labels(cm, "it'll") == [60, 45, 1, 79]          # test_dataProcessor.py:139-143
    assert pkg.get_str_labels(cm, "'d") == [0, 79]                      # :145-149
    ids = pkg.get_str_labels(cm, "i will")                              # :191-229  "IWill_", one-hot columns
    assert ids == [60, 74, 34, 16, 79] and [cm[i] for i in ids] == ["I", "W", "i", "ll", "_"]
    onehot_full = pkg.get_str_to_one_hot_encoded(cm, "i will")
    assert [int(v.argmax()) for v in onehot_full] == [60, 74, 34, 16, 79] and all(v.sum() == 1 for v in onehot_full)
    for text in ("hello world", "it's a test", "we've seen mississippi"):
        assert pkg.get_labels_str(cm, pkg.get_str_labels(cm, text)) == text
    onehot = pkg.get_str_to_one_hot_encoded(cm, "ab", add_eos=True)
    assert len(onehot) == 3 and onehot[0].argmax() == cm.index("A") and onehot[2].argmax() == 79


def test_matches_reference_codec_golden(pkg):
    g = golden("labels.npz")
    cm = [str(c) for c in g["char_map"]]
    assert cm == pkg.ENGLISH_CHAR_MAP
    for i, text in enumerate(g["texts"]):
        ids = pkg.get_str_labels(cm, str(text))
        np.testing.assert_array_equal(ids, g["ids_%d" % i])
        assert pkg.get_labels_str(cm, ids) == str(g["back_%d" % i])


def test_get_labels_str_drops_out_of_range_ids(pkg):
    cm = pkg.ENGLISH_CHAR_MAP
    assert pkg.get_labels_str(cm, [59, 30, 80, 80, -1]) == "he"       # 80 = process_input padding
