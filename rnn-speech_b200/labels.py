# coding=utf-8
"""Label codec and character map (host side, pure Python).

Same behaviour as DataProcessor.get_str_labels / get_labels_str /
get_str_to_one_hot_encoded (/root/reference/util/dataprocessor.py:97-205) and
the 80-symbol map of /root/reference/models/SpeechRecognizer.py:21-36; pinned by
the reference's known-answer tests (util/test_dataProcessor.py:132-229), which
tests/test_labels.py restates.
"""
import logging

import numpy as np

_APOSTROPHES = ["'d", "'ll", "'m", "'nt", "'s", "s'", "'t", "'ve"]
_DOUBLES = [c + c for c in "bcdefgiklmnoprstuz"]
_LOWER = [chr(c) for c in range(ord("a"), ord("z") + 1)]
_UPPER = [c.upper() for c in _LOWER]
ENGLISH_CHAR_MAP = _APOSTROPHES + _DOUBLES + _LOWER + _UPPER + ["'", "_"]
assert len(ENGLISH_CHAR_MAP) == 80


def capitalise_word_starts(text):
    """Drop spaces; the first letter of every word becomes a capital."""
    out, start = [], True
    for ch in text:
        if ch == " ":
            start = True
        elif start:
            out.append(ch.upper())
            start = False
        else:
            out.append(ch)
    return "".join(out)


def get_str_labels(char_map, _str, add_eos=True):
    """String -> label ids: greedy longest match (3, 2 then 1 characters; the
    multi-character tokens are matched case-insensitively, single characters
    case-sensitively), EOS = len(char_map) - 1 appended."""
    index = {tok: i for i, tok in reversed(list(enumerate(char_map)))}
    text = capitalise_word_starts(_str)
    ids, pos = [], 0
    while pos < len(text):
        for width in (3, 2, 1):
            if len(text) - pos < width:
                continue
            piece = text[pos:pos + width]
            key = piece.lower() if width > 1 else piece
            if key in index:
                ids.append(index[key])
                pos += width
                break
        else:
            logging.warning("Unable to process label : %s", text)
            break
    if add_eos:
        ids.append(len(char_map) - 1)
    return ids


def get_labels_str(char_map, label):
    """Label ids -> text: ids outside the map are dropped, the FIRST EOS token is
    removed, a space goes before every capital except the first character."""
    chars = [char_map[i] for i in label if 0 <= i < len(char_map)]
    if char_map[-1] in chars:
        chars.remove(char_map[-1])
    words = []
    for n, tok in enumerate(chars):
        if n != 0 and tok.isupper():
            words.append(" ")
        words.append(tok.lower())
    return "".join(words)


def get_str_to_one_hot_encoded(char_map, _str, add_eos=True):
    ids = get_str_labels(char_map, _str, add_eos=add_eos)
    out = []
    for i in ids:
        v = np.zeros(len(char_map))
        v[i] = 1
        out.append(v)
    return out
