"""ORACLE — TEST INFRASTRUCTURE ONLY.  Shared glue that turns one scalar-type
submodule of ``_gtn_oracle`` into the public ``gtn`` namespace (signature quirks
of the reference's call sites are absorbed here, not in C++)."""

CPU = 0
CUDA = 1


class Device:
    """gtn.Device(gtn.CPU) — the reference always selects the CPU device
    (ctc.py:41, asg.py:97,219, stc.py:75, transducer.py:212,262,487)."""

    def __init__(self, kind=CPU, index=0):
        self.kind = kind
        self.index = index

    def __repr__(self):
        return "Device(%s)" % ("CPU" if self.kind == CPU else "CUDA")


def bind(ns, ext):
    for name in dir(ext):
        if not name.startswith("_"):
            ns[name] = getattr(ext, name)

    def linear_graph(M, N, *args, **kwargs):
        """linear_graph(M, N, device, calc_grad) (ctc.py:40) and the older
        linear_graph(M, N, calc_grad).  tests/transducer_test.py:256 passes a
        Graph where a Device is meant; any non-bool third argument is taken
        as the device and ignored."""
        calc_grad = kwargs.get("calc_grad", True)
        if len(args) == 1:
            if isinstance(args[0], bool):
                calc_grad = args[0]
        elif len(args) >= 2:
            calc_grad = args[1]
        return ext.linear_graph_(M, N, bool(calc_grad))

    def write_dot(*args, **kwargs):  # debugging aid in the reference; no-op here
        return None

    ns["linear_graph"] = linear_graph
    ns["write_dot"] = write_dot
    ns["draw"] = write_dot
    ns["Device"] = Device
    ns["CPU"] = CPU
    ns["CUDA"] = CUDA
