"""Route the Python package to the host-emulated library (TEST INFRASTRUCTURE, see cuda_emu.h).

`install()` must run before `sph_taichi_b200._lib` is imported: it points SPH_B200_LIB at
tests/emu/_build/libsph_b200_emu*.so, swaps `engine.Engine` for a subclass that keeps its workspace in host
memory, and makes `ParticleSystem` default to the CPU device.  Nothing in the package itself changes."""
import ctypes as C
import os


def install(lib_path):
    os.environ["SPH_B200_LIB"] = lib_path
    import torch
    from sph_taichi_b200 import _lib, engine, particle_system

    assert _lib.LIB_PATH == lib_path, "sph_taichi_b200._lib was imported before emu_engine.install()"

    class EmuEngine(engine.Engine):
        def __init__(self, params, n_max, n_solid=0, n_bodies=0, device=None):
            self.lib = _lib.load()
            self.device = torch.device("cpu")
            self.params = params
            self.n_max = int(n_max)
            self.n_solid_cap = int(n_solid)
            nbytes = self.lib.sph_workspace_bytes(C.byref(params), self.n_max, self.n_solid_cap, int(n_bodies))
            if nbytes == 0:
                raise ValueError("invalid SPH parameters (grid must be >= 3 cells per axis, dim == 3)")
            # uninitialised on purpose, like device memory (0xAB pattern makes stale reads visible)
            self.workspace = torch.full((int(nbytes) + 256,), 0xAB, dtype=torch.uint8)
            self._ws_ptr = (self.workspace.data_ptr() + 255) // 256 * 256
            handle = C.c_void_p()
            rc = self.lib.sph_create(C.byref(params), self.n_max, self.n_solid_cap, int(n_bodies), 0,
                                     C.c_void_p(self._ws_ptr), int(nbytes), C.byref(handle))
            if rc:
                raise RuntimeError(f"sph_create failed ({rc}): {self.lib.sph_last_error(None).decode()}")
            self.ctx = handle

        def _stream(self):
            return C.c_void_p(0)

        def close(self):
            if getattr(self, "ctx", None):
                self.lib.sph_destroy(self.ctx)
                self.ctx = None

    engine.Engine = EmuEngine
    torch.cuda.is_available = lambda: True

    # the emulated runtime is synchronous: streams and events of the slab driver become no-ops
    class _Stream:
        cuda_stream = 0

        def __init__(self, *a, **k):
            pass

        def wait_stream(self, *a):
            pass

        def wait_event(self, *a):
            pass

        def synchronize(self):
            pass

    class _Event:
        def __init__(self, *a, **k):
            pass

        def record(self, *a):
            pass

        def wait(self, *a):
            pass

        def synchronize(self):
            pass

        def elapsed_time(self, other):
            return 0.0

    import contextlib
    torch.cuda.synchronize = lambda *a, **k: None
    torch.cuda.set_device = lambda *a, **k: None
    torch.cuda.current_stream = lambda *a, **k: _Stream()
    torch.cuda.Stream = _Stream
    torch.cuda.Event = _Event
    torch.cuda.stream = lambda s: contextlib.nullcontext()
    _init = particle_system.ParticleSystem.__init__

    def init_on_cpu(self, config, GGUI=False, device=None):
        _init(self, config, GGUI=GGUI, device="cpu")

    particle_system.ParticleSystem.__init__ = init_on_cpu
    return EmuEngine
