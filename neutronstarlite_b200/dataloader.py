"""`GNNDatum` (core/ntsDataloador.hpp:27-330): the feature / label / mask tables of the vertices a rank owns.

Same three entry points the reference's toolkits call (`random_generate`, `readFeature_Label_Mask`, `registLabel` /
`registMask`), backed by the host routines of libnts_b200 (`nts_host_read_feature_label_mask`: the reference's text
tables parsed in parallel; `nts_host_read_feature_binary`: a packed float32 table, one pread of the owned rows)."""
from __future__ import annotations

import numpy as np

from . import _lib


class GNNDatum:
    def __init__(self, feature_size, label_num, v_begin, v_end):
        self.feature_size = int(feature_size)
        self.label_num = int(label_num)
        self.p_v_s, self.p_v_e = int(v_begin), int(v_end)
        n = self.p_v_e - self.p_v_s
        self.local_feature = np.zeros((n, self.feature_size), dtype=np.float32)
        self.local_label = np.zeros(n, dtype=np.int64)
        # the reference memsets the int mask with byte 1 (0x01010101): rows no table mentions keep that value
        self.local_mask = np.full(n, 0x01010101, dtype=np.int32)

    def random_generate(self, seed=0):
        """core/ntsDataloador.hpp:63-71: all-ones features, labels rand() % label_num, mask v % 3 (libc rand() is
        replaced by a seeded numpy generator)."""
        n = self.p_v_e - self.p_v_s
        self.local_feature[:] = 1.0
        self.local_label[:] = np.random.default_rng(seed).integers(0, self.label_num, n)
        self.local_mask[:] = np.arange(n) % 3

    def readFeature_Label_Mask(self, inputF, inputL, inputM):
        """core/ntsDataloador.hpp:156-221."""
        rc = _lib.load().nts_host_read_feature_label_mask(
            str(inputF).encode(), str(inputL).encode() if inputL else None, str(inputM).encode() if inputM else None,
            self.feature_size, self.p_v_s, self.p_v_e, self.local_feature.ctypes.data,
            self.local_label.ctypes.data if inputL else None, self.local_mask.ctypes.data if inputM else None)
        if rc != 0:
            raise _lib.NtsError("nts_host_read_feature_label_mask failed (rc=%d): unreadable or malformed table" % rc)
        return self

    def read_feature_binary(self, path):
        """Packed float32 [V, feature_size] table: only the owned rows are read."""
        rc = _lib.load().nts_host_read_feature_binary(str(path).encode(), self.feature_size, self.p_v_s, self.p_v_e,
                                                      self.local_feature.ctypes.data)
        if rc != 0:
            raise _lib.NtsError("nts_host_read_feature_binary failed (rc=%d)" % rc)
        return self

    # registLabel / registMask / the feature tensor (core/ntsDataloador.hpp:78-92): torch views of the host arrays
    def registLabel(self, device=None):
        import torch
        t = torch.from_numpy(self.local_label)
        return t.to(device) if device is not None else t

    def registMask(self, device=None):
        import torch
        t = torch.from_numpy(self.local_mask).view(-1, 1)
        return t.to(device) if device is not None else t

    def features(self, device=None, pinned=False):
        import torch
        t = torch.from_numpy(self.local_feature)
        if pinned:
            t = t.pin_memory()
        return t.to(device, non_blocking=pinned) if device is not None else t
