"""Drop-in for the N-run splitter of dbb/splitScaffolds.py (class ``SplitScaffolds``, method ``splits``), GPU backed.

``splits(scaffoldId, scaffold, minN)`` keeps the reference's signature and return value (splitScaffolds.py:27-53):
a dict ``'<scaffoldId>_c<k>' -> [start, end]``.  ``splitsBatch`` is the array-native form the batched preprocess path
uses (one launch for all scaffolds).  Reference quirks are kept, see csrc/split.cu.
"""
import logging

import numpy as np


class SplitScaffolds(object):
    def __init__(self):
        self.logger = logging.getLogger()

    def splitsBatch(self, scaffoldIds, scaffolds, minN):
        """scaffolds: list of str/bytes.  Returns a list of dicts, one per scaffold, each as ``splits`` returns."""
        import torch
        from . import device
        raw = [s.encode('latin-1') if isinstance(s, str) else bytes(s) for s in scaffolds]
        off = np.zeros(len(raw) + 1, dtype=np.int64)
        off[1:] = np.cumsum([len(r) for r in raw])
        flat = np.frombuffer(b''.join(raw) + b'\0\0\0\0', dtype=np.uint8)
        device.require_cuda()
        d_ascii = torch.from_numpy(flat.copy()).cuda()
        packed = device.pack_ranges(d_ascii, off[:-1], off[1:], want_codes=False, want_nmask=True)
        c_off, begin, end = device.split_scaffolds(packed, minN)
        out = []
        for i, sid in enumerate(scaffoldIds):
            d = {}
            for k in range(int(c_off[i]), int(c_off[i + 1])):
                d[sid + '_c' + str(k - int(c_off[i]))] = [int(begin[k]), int(end[k])]
            out.append(d)
        return out

    def splits(self, scaffoldId, scaffold, minN):
        return self.splitsBatch([scaffoldId], [scaffold], minN)[0]
