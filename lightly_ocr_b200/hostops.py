"""Host-side logic of the path that the reference also keeps on the host: the reading-order comparator, the label
converters' decode step, and crop slicing.  Pure Python / numpy, no GPU, no oracle imports."""
from functools import cmp_to_key

ALPHABET = "0123456789abcdefghijklmnopqrstuvwxyz"


def compare_rects(first_rect, second_rect):
    """Reading-order comparator with the reference's exact branch structure (ocr/tools/det_utils.py:8-26), including
    the two branches that compare a rect with itself.  It is not a consistent total order: the sorted result is
    whatever CPython's list sort makes of it, which is why the sort stays in Python."""
    a, b = first_rect, second_rect
    if a[2] <= b[0]:
        return -1
    if b[2] <= a[0]:
        return 1
    if a[3] <= a[1]:
        return -1
    if b[2] <= b[0]:
        return 1
    for i in (1, 0, 3, 2):
        if a[i] != b[i]:
            return -1 if a[i] < b[i] else 1
    return 0


def sort_rects(rects):
    """sorted(rects, key=cmp_to_key(compare_rects))  (reference ocr/net.py:108)."""
    return sorted(rects, key=cmp_to_key(compare_rects))


class CTCLabelConverter:
    """decode() of the reference's converter (ocr/tools/recog_utils.py:10-47): index 0 is the CTC blank."""

    def __init__(self, character):
        self.dict = {c: i + 1 for i, c in enumerate(character)}
        self.character = ["[blank]"] + list(character)

    def encode(self, text, batch_max_len=25):
        """(concatenated class indices int32, lengths int32) of a list of labels (recog_utils.py:24-30)."""
        import numpy as np
        length = [len(t) for t in text]
        return (np.array([self.dict[c] for c in "".join(text)], np.int32), np.array(length, np.int32))

    def decode(self, text, length):
        texts, index = [], 0
        for l in length:
            l = int(l)
            t = text[index:index + l]
            chars = [self.character[int(t[i])] for i in range(l)
                     if int(t[i]) != 0 and not (i > 0 and int(t[i - 1]) == int(t[i]))]
            texts.append("".join(chars))
            index += l
        return texts


class AttnLabelConverter:
    """decode() of the reference's attention converter (recog_utils.py:50-119): '[GO]' = 0, '[s]' = 1."""

    def __init__(self, character):
        self.character = ["[GO]", "[s]"] + list(character)
        self.dict = {c: i for i, c in enumerate(self.character)}

    def encode(self, text, batch_max_len=25):
        """([n, batch_max_len + 2] int32 rows: [GO], the label's tokens, [s], [GO] padding; lengths + 1) of a list of
        labels (recog_utils.py:84-96).  The reference returns from inside its loop, i.e. fills row 0 only; every row is
        filled here."""
        import numpy as np
        length = [len(t) + 1 for t in text]
        out = np.zeros((len(text), batch_max_len + 2), np.int32)
        for i, t in enumerate(text):
            ids = [self.dict[c] for c in t] + [self.dict["[s]"]]
            out[i, 1:1 + len(ids)] = ids
        return out, np.array(length, np.int32)

    def decode(self, text, length):
        return ["".join(self.character[int(i)] for i in text[index]) for index, _ in enumerate(length)]
