"""TEST INFRASTRUCTURE: an oracle-backed stand-in for ntsc_crt_b200.capi.Batch (same methods, CPU torch tensors),
so that the host-side scheduling built on the batch interface -- video.VideoConverter: segments, halos,
speculation, verification, repair, the exchanges between ranks -- can run in the CPU suite, incl. over gloo.
Injected through VideoConverter(batch_factory=...); the product never imports this."""
import support as S
from ntsc_crt_b200 import capi, layout


class OracleBatch:
    def __init__(self, variant, n):
        self.variant, self.n = variant, n
        self.spec = layout.system_spec(variant)
        self._mon = [None] * n      # (out tensor, fmt, noise, knobs)
        self._eng = [None] * n
        self._src = [None] * n      # (image tensor, settings)
        self.launches = 0

    def set_option(self, name, value):
        pass

    def set_monitor(self, i, out, fmt=layout.PIX_BGRA, noise=0, **knobs):
        self._mon[i] = (out, fmt, noise, knobs)

    def commit_monitors(self, first=0, count=None):
        count = self.n - first if count is None else count
        for i in range(first, first + count):
            out, fmt, noise, knobs = self._mon[i]
            eng = S.OracleEngine(self.variant, out.shape[1], out.shape[0], fmt, out=out.numpy())  # shares memory
            eng.set(**knobs)
            self._eng[i] = eng

    def set_source(self, i, img, **settings):
        self._src[i] = (img, dict(settings))

    def modulate(self, stream=0, first=0, count=None):
        count = self.n - first if count is None else count
        for i in range(first, first + count):
            img, settings = self._src[i]
            self._eng[i].modulate(img.numpy(), **settings)
        self.launches += 3

    def demodulate(self, stream=0, first=0, count=None):
        count = self.n - first if count is None else count
        for i in range(first, first + count):
            self._eng[i].demodulate(self._mon[i][2])
        self.launches += 3

    def get_state(self, first=0, count=None, stream=0):
        count = self.n - first if count is None else count
        st = (capi.State * count)()
        for k in range(count):
            m = self._eng[first + k].mon
            st[k].hsync, st[k].vsync, st[k].rn = m.hsync, m.vsync, m.rn
            for r in range(3):
                for x in range(4):
                    st[k].ccf[r][x] = m.ccf[r][x]
        return st

    def set_state(self, states, first=0, stream=0):
        for k in range(len(states)):
            m = self._eng[first + k].mon
            m.hsync, m.vsync, m.rn = states[k].hsync, states[k].vsync, states[k].rn
            for r in range(3):
                for x in range(4):
                    m.ccf[r][x] = states[k].ccf[r][x]

    def close(self):
        self._eng = [None] * self.n
