"""MT19937 restatement (oracle; test infrastructure only).

Follows packages/basics/random/c_src/MersenneTwister.cc of the reference:
  initialize()  :226-240   (Knuth 1812433253 seeding)
  reload()      :243-255   (twist over 624 words)
  randInt()     :104-116   (tempering)
  randInt(n)    :118-134   (mask + rejection)
  rand(n)       :62-68     (randInt * 1/4294967295 * n)
  shuffle()     :279-288   (Fisher-Yates from the back, randInt(i))
The generator is the published MT19937 of Matsumoto & Nishimura.
"""
import numpy as np

_N, _M = 624, 397


class MTRand:
    def __init__(self, seed):
        self.seed(seed)

    # MersenneTwister.cc:137-141 + :226-240
    def seed(self, one_seed):
        s = np.empty(_N, dtype=np.uint64)
        s[0] = np.uint64(int(one_seed) & 0xFFFFFFFF)
        for i in range(1, _N):
            prev = int(s[i - 1])
            s[i] = (1812433253 * (prev ^ (prev >> 30)) + i) & 0xFFFFFFFF
        self.state = s.astype(np.uint32)
        self._reload()

    # MersenneTwister.cc:243-255 ; twist() MersenneTwister.h:141-142
    def _reload(self):
        st = [int(v) for v in self.state]

        def twist(m, s0, s1):
            mix = (s0 & 0x80000000) | (s1 & 0x7FFFFFFF)
            return m ^ (mix >> 1) ^ (0x9908B0DF if (s1 & 1) else 0)

        for i in range(_N - _M):
            st[i] = twist(st[i + _M], st[i], st[i + 1])
        for i in range(_N - _M, _N - 1):
            st[i] = twist(st[i + _M - _N], st[i], st[i + 1])
        st[_N - 1] = twist(st[_M - 1], st[_N - 1], st[0])
        self.state = np.array(st, dtype=np.uint32)
        # temper the whole block at once (MersenneTwister.cc:110-115)
        y = self.state.copy()
        y ^= y >> np.uint32(11)
        y ^= (y << np.uint32(7)) & np.uint32(0x9D2C5680)
        y ^= (y << np.uint32(15)) & np.uint32(0xEFC60000)
        y ^= y >> np.uint32(18)
        self._out = y
        self._pos = 0

    def randInt32(self):
        if self._pos >= _N:
            self._reload()
        v = int(self._out[self._pos])
        self._pos += 1
        return v

    def raw(self, n):
        """n consecutive raw 32-bit outputs as a uint32 array."""
        out = np.empty(n, dtype=np.uint32)
        k = 0
        while k < n:
            if self._pos >= _N:
                self._reload()
            take = min(n - k, _N - self._pos)
            out[k:k + take] = self._out[self._pos:self._pos + take]
            self._pos += take
            k += take
        return out

    # MersenneTwister.cc:118-134
    def randInt(self, n=None, hi=None):
        """randInt() -> [0,2^32-1]; randInt(n) -> [0,n]; randInt(lo,hi) -> [lo,hi]
        (the two-argument form is the Lua binding, bind_mtrand.lua.cc:157-166)."""
        if n is None:
            return self.randInt32()
        if hi is not None:
            return n + self.randInt(hi - n)
        used = n
        used |= used >> 1
        used |= used >> 2
        used |= used >> 4
        used |= used >> 8
        used |= used >> 16
        while True:
            i = self.randInt32() & used
            if i <= n:
                return i

    # MersenneTwister.cc:62-68
    def rand(self, n=1.0):
        return float(self.randInt32()) * (1.0 / 4294967295.0) * n

    def rand_array(self, count, n=1.0):
        """count consecutive rand(n) draws as float64."""
        return self.raw(count).astype(np.float64) * (1.0 / 4294967295.0) * n

    # MersenneTwister.cc:279-288 ; the Lua binding returns 1-based indices,
    # here they stay 0-based.
    def shuffle(self, size):
        v = list(range(size))
        for i in range(size - 1, 0, -1):
            j = self.randInt(i)
            v[i], v[j] = v[j], v[i]
        return v
