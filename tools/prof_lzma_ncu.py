import sys, lzma
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import helpers as H
from swcompression_b200 import XZArchive
raw = H.textlike(1 << 20, 5)
comp = lzma.compress(raw, format=lzma.FORMAT_XZ, check=lzma.CHECK_CRC64, filters=[{"id": lzma.FILTER_LZMA2, "preset": 6, "dict_size": 1 << 20}])
assert XZArchive.unarchive(comp) == raw
