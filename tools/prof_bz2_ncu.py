import sys, bz2
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import helpers as H
from swcompression_b200 import BZip2
raw = H.textlike(900000, 4)
comp = bz2.compress(raw, 9)
assert BZip2.decompress(comp) == raw
