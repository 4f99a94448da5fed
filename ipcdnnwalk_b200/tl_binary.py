"""Readers for taesooLib's on-disk formats met on the path.

  BinaryFile  (work/taesooLib/BaseLib/utility/tfile.h:123, tfile.cpp:158-420): little-endian stream of
              (int32 type tag, payload): INT, FLOAT (double), FLOATN / SPFLOATN (n, values), INTN, BITN, FLOATMN /
              SPFLOATMN (rows, cols, values), STRING (len, bytes), ...
  util.saveTable / SaveTable:pickle_ (work/taesooLib/Resource/scripts/module.lua:8036-8260): a Lua table graph written
              as  numRef, then for ref = numRef .. 1:  { 1, isRef, key, value | ref } ... 0   with keys / scalar values
              pickled as (kind, payload): 0 number, 1 string, 2 bool, 3 vector3, 4 quater, 5 vectorn / matrixn, ...
Used for gym_cdm2/spec/refTraj_*.dat (recorded simulator roll-outs of the AdaptiveSRB environment)."""
import struct
import numpy as np

T_INT, T_FLOAT, T_FLOATN, T_INTN, T_BITN, T_FLOATMN, T_INTMN, T_BITMN, T_STRING, T_STRINGN, T_ARRAY, T_EOF, T_SPFLOAT, T_SPFLOATN, \
    T_SPFLOATMN, T_FLOATND, T_SPFLOATND = range(17)


class BinaryReader:
    def __init__(self, data):
        self.b = data
        self.p = 0

    def _int(self):
        v = struct.unpack_from("<i", self.b, self.p)[0]
        self.p += 4
        return v

    def _arr(self, dtype, n):
        a = np.frombuffer(self.b, dtype=dtype, count=n, offset=self.p)
        self.p += a.nbytes
        return a

    def eof(self):
        return self.p >= len(self.b)

    def unpack_int(self):
        t = self._int()
        if t != T_INT:
            raise ValueError(f"unpackInt: tag {t} at {self.p - 4}")
        return self._int()

    def unpack_any(self):
        """One tagged value -> python / numpy object (BinaryFile::unpackAny)."""
        t = self._int()
        if t == T_INT:
            return self._int()
        if t == T_FLOAT:
            return float(self._arr("<f8", 1)[0])
        if t == T_SPFLOAT:
            return float(self._arr("<f4", 1)[0])
        if t == T_FLOATN:
            return self._arr("<f8", self._int()).copy()
        if t == T_SPFLOATN:
            return self._arr("<f4", self._int()).astype(np.float64)
        if t == T_INTN:
            return self._arr("<i4", self._int()).copy()
        if t == T_FLOATMN:
            r, c = self._int(), self._int()
            return self._arr("<f8", r * c).reshape(r, c).copy()
        if t == T_SPFLOATMN:
            r, c = self._int(), self._int()
            return self._arr("<f4", r * c).reshape(r, c).astype(np.float64)
        if t == T_INTMN:
            r, c = self._int(), self._int()
            return self._arr("<i4", r * c).reshape(r, c).copy()
        if t == T_BITN:
            n = self._int()
            nbytes = (n + 7) // 8
            bits = np.unpackbits(self._arr("u1", nbytes), bitorder="little")[:n]
            return bits.astype(bool)
        if t == T_STRING:
            n = self._int()
            s = bytes(self.b[self.p:self.p + n]).decode("latin1")
            self.p += n
            return s
        if t == T_ARRAY:
            count, size = self._int(), self._int()
            return self._arr("u1", count * size).copy()
        raise ValueError(f"unsupported BinaryFile tag {t} at byte {self.p - 4}")

    def unpack_pickle(self):
        kind = self.unpack_int()
        if kind == 0:
            return self.unpack_any()                # number
        if kind == 1:
            return self.unpack_any()                # string
        if kind == 2:
            return self.unpack_int() == 1           # bool
        if kind == 9:                               # transf: translation, rotation
            return (self.unpack_any(), self.unpack_any())
        if kind == -1:
            raise ValueError("pickled userdata with a custom pack method is not supported")
        return self.unpack_any()                    # vector3 / quater (x y z w) / vectorn / matrixn / quaterN / vector3N / boolN


def load_table(path):
    """util.loadTable: -> nested dict (Lua tables; integer-valued number keys become ints)."""
    r = BinaryReader(open(path, "rb").read())
    num_ref = r.unpack_int()
    tables = {i: {} for i in range(1, num_ref + 1)}
    for i in range(num_ref, 0, -1):
        t = tables[i]
        while r.unpack_int() == 1:
            is_ref = r.unpack_int()
            k = r.unpack_pickle()
            if isinstance(k, float) and k == int(k):
                k = int(k)
            t[k] = tables[r.unpack_int()] if is_ref == 1 else r.unpack_pickle()
    return tables[1]
