"""oracle/codec.py - TEST INFRASTRUCTURE. Pure-Python restatement of the string side of
utils/ctc_codec.py: the class table (:17-30), encode (:43-61) and index->text mapping (:94-95)."""
import numpy as np


class CodecTables(object):
    def __init__(self, characters_str):
        self.chars_list = list(characters_str)                                   # :19
        self.dict = {}
        for i, ch in enumerate(self.chars_list):                                 # :22-24 (duplicates: last wins)
            self.dict[ch] = i + 1
        self.characters = ['<blank>'] + self.chars_list + ['<unknown>']         # :28
        self.dict['<blank>'] = 0
        self.dict['<unknown>'] = len(self.characters) - 1

    def encode(self, text):
        length = [len(s) for s in text]                                          # :51
        index = []
        for ch in ''.join(text):                                                 # :52-59
            if ch in self.chars_list:
                index.append(self.dict[ch])
            else:
                index.append(len(self.characters) - 1)
        return np.array(index, dtype=np.int32), np.array(length, dtype=np.int32)

    def to_text(self, idx, length):
        return [''.join(self.characters[i] for i in idx[b, :length[b]]) for b in range(len(length))]
