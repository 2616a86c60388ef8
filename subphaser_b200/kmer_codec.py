"""Host-side k-mer text <-> 2-bit key conversion (wire-format glue, vectorised numpy).
Key encoding = libspk's: 2 bits per base, first base most significant, A=0 C=1 G=2 T=3."""
import numpy as np

_ALPHA = np.frombuffer(b"ACGT", dtype=np.uint8)
_CODE = np.full(256, 255, dtype=np.uint8)
for _i, _c in enumerate(b"ACGT"):
    _CODE[_c] = _i
    _CODE[ord(chr(_c).lower())] = _i


def keys_to_bytes(keys, k):
    """uint64 [M] -> uint8 [M, k] ASCII"""
    keys = np.asarray(keys, dtype=np.uint64)
    shifts = (2 * (k - 1 - np.arange(k))).astype(np.uint64)
    codes = ((keys[:, None] >> shifts[None, :]) & np.uint64(3)).astype(np.uint8)
    return _ALPHA[codes]


def keys_to_strs(keys, k):
    b = keys_to_bytes(keys, k)
    if b.shape[0] == 0:
        return []
    return b.view("S%d" % k).ravel().astype("U%d" % k).tolist()


def strs_to_keys(strs, k=None):
    """iterable of str -> (uint64 keys, valid mask).  Strings with non-ACGT characters are invalid."""
    strs = list(strs)
    if not strs:
        return np.zeros(0, np.uint64), np.zeros(0, bool)
    if k is None:
        k = len(strs[0])
    arr = np.array(strs, dtype="S%d" % k)
    b = np.frombuffer(arr.tobytes(), dtype=np.uint8).reshape(len(strs), k)
    codes = _CODE[b]
    valid = (codes != 255).all(axis=1) & (np.char.str_len(arr) == k)
    codes = np.where(codes == 255, 0, codes).astype(np.uint64)
    shifts = (2 * (k - 1 - np.arange(k))).astype(np.uint64)
    keys = (codes << shifts[None, :]).sum(axis=1, dtype=np.uint64)
    return keys, valid


def revcomp_keys(keys, k):
    keys = np.asarray(keys, dtype=np.uint64)
    out = np.zeros_like(keys)
    x = keys.copy()
    for _ in range(k):
        out = (out << np.uint64(2)) | (np.uint64(3) - (x & np.uint64(3)))
        x = x >> np.uint64(2)
    return out


def canonical_keys(keys, k):
    rc = revcomp_keys(keys, k)
    return np.minimum(np.asarray(keys, dtype=np.uint64), rc)


def revcomp_str(s):
    return s.translate(str.maketrans("ACGTacgt", "TGCAtgca"))[::-1]
