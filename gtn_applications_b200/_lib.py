"""ctypes binding of libwfst_b200.so (C ABI: include/wfst_b200.h).

The library is built in-tree (gtn_applications_b200/lib/) by
``__graft_entry__.build()`` / ``make -C gtn_applications_b200/csrc``.  Loading
fails loudly: there is no fallback implementation."""
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("WFST_B200_LIB") or os.path.join(_HERE, "lib", "libwfst_b200.so")

c_float_p = ctypes.c_void_p
c_int_p = ctypes.c_void_p


class AcceptorBatch(ctypes.Structure):
    """wfst_acceptor_batch_t"""
    _fields_ = [
        ("B", ctypes.c_int32), ("max_nodes", ctypes.c_int32), ("max_arcs", ctypes.c_int32),
        ("node_offsets", ctypes.c_void_p), ("arc_offsets", ctypes.c_void_p),
        ("node_flags", ctypes.c_void_p),
        ("in_ptr", ctypes.c_void_p), ("in_src", ctypes.c_void_p), ("in_label", ctypes.c_void_p),
        ("in_arc", ctypes.c_void_p),
        ("out_ptr", ctypes.c_void_p), ("out_dst", ctypes.c_void_p), ("out_label", ctypes.c_void_p),
        ("out_arc", ctypes.c_void_p),
        ("weights", ctypes.c_void_p),
        ("final_weights", ctypes.c_void_p), ("grad_final_weights", ctypes.c_void_p),
    ]


# name -> (restype, argtypes); the single source of truth for the symbols the
# library must export (tests/test_capi_symbols.py checks it against the header)
_I, _Z, _P, _I32 = ctypes.c_int, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_int32
SIGNATURES = {
    "wfst_last_error": (ctypes.c_char_p, []),
    "wfst_abi_version": (_I, []),
    "wfst_launch_count": (ctypes.c_ulonglong, []),
    "wfst_debug_force_generic_ctc": (_I, [_I]),
    "wfst_debug_ctc_chain_config": (_I, [_I, _I]),
    "wfst_debug_force_generic_lattice": (_I, [_I]),
    "wfst_asg_viterbi_supported": (_I, [_I, _I]),
    "wfst_asg_viterbi": (_I, [_P, _P, _I, _I, _I, _P, _P, _P]),
    "wfst_debug_ctc_hazards": (_I, [_P, _I, _I, _I, _I, _P]),
    "wfst_ctc_workspace_bytes": (_Z, [_I, _I, _I, _I]),
    "wfst_ctc_forward_backward": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _Z, _P]),
    "wfst_ctc_logits_supported": (_I, [_I, _I, _I, _I]),
    "wfst_ctc_logits_workspace_bytes": (_Z, [_I, _I, _I, _I]),
    "wfst_ctc_logits_forward_backward": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _Z, _P]),
    "wfst_ctc_forward_backward_host": (_I, [_P, _P, _P, _I, _I, _I, _I, _P, _P, _P, _P]),
    "wfst_lattice_workspace_bytes": (_Z, [_I, _I, _I, _I, _I]),
    "wfst_lattice_forward_backward": (_I, [_P, _I, _I, _I, ctypes.POINTER(AcceptorBatch), _I, _P, _P,
                                           _P, _I, _P, _P, _Z, _P]),
    "wfst_lattice_forward_backward_cross": (_I, [_P, _I, _I, _I, _P, _P, _P, _P, _P, _P, _Z, _P]),
    "wfst_lattice_forward_backward_many": (_I, [_P, _I, _I, _I, _P, _I, _P, _P, _P, _P, _P, _Z, _P]),
    "wfst_asg_workspace_bytes": (_Z, [_I, _I, _I, _I]),
    "wfst_asg_forward_backward": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _Z, _P]),
    "wfst_scale_inplace": (_I, [_P, _Z, _P, _P]),
    # host-side graphs (csrc/graph.cpp)
    "wfst_graph_create": (_I32, [_I]),
    "wfst_graph_destroy": (_I, [_I32]),
    "wfst_graph_destroy_many": (_I, [_P, _I]),
    "wfst_graph_add_node": (_I, [_I32, _I, _I]),
    "wfst_graph_add_arc": (_I, [_I32, _I, _I, _I, _I, ctypes.c_float]),
    "wfst_graph_add_arcs": (_I, [_I32, _I, _P, _P, _P, _P, _P]),
    "wfst_graph_num_nodes": (_I, [_I32]),
    "wfst_graph_num_arcs": (_I, [_I32]),
    "wfst_graph_arc_sort": (_I, [_I32, _I]),
    "wfst_graph_mark_arc_sorted": (_I, [_I32, _I]),
    "wfst_graph_sorted_flags": (_I, [_I32]),
    "wfst_graph_get_calc_grad": (_I, [_I32]),
    "wfst_graph_set_calc_grad": (_I, [_I32, _I]),
    "wfst_graph_set_weights": (_I, [_I32, _P]),
    "wfst_graph_get_weights": (_I, [_I32, _P]),
    "wfst_graph_get_arcs": (_I, [_I32, _P, _P, _P, _P]),
    "wfst_graph_get_node_flags": (_I, [_I32, _P]),
    "wfst_graph_get_arc_order": (_I, [_I32, _I, _P]),
    "wfst_graph_get_provenance": (_I, [_I32, _P, _P]),
    "wfst_graph_compose": (_I32, [_I32, _I32]),
    "wfst_graph_remove": (_I32, [_I32, _I, _I]),
    "wfst_graph_project": (_I32, [_I32, _I]),
    "wfst_graph_linear": (_I32, [_I, _I, _I]),
    "wfst_graph_loadtxt": (_I32, [ctypes.c_char_p]),
    "wfst_graph_savetxt": (_I, [_I32, ctypes.c_char_p]),
    "wfst_graph_load": (_I32, [ctypes.c_char_p]),
    "wfst_graph_save": (_I, [_I32, ctypes.c_char_p]),
    "wfst_stc_graphs": (_I, [_P, _P, _I, _I, ctypes.c_float, _I, _P]),
    "wfst_transducer_alignment_graphs": (_I, [_I32, _I32, _P, _P, _I, _P]),
    "wfst_transducer_decode_paths": (_I, [_I32, _P, _I, _I, _P, _P]),
    "wfst_graph_viterbi_path": (_I32, [_I32]),
    "wfst_graph_score": (_I, [_I32, _I, _P, _P]),
    "wfst_transducer_alignment_cache": (_I, [ctypes.c_longlong, _P, _P]),
    "wfst_fold_transitions_batch": (_I32, [_I32, _P, _I, _P]),
    "wfst_fold_sizes": (_I, [_I32, _P]),
    "wfst_fold_fill": (_I, [_I32, _P, _P, _P, _P, _P]),
    "wfst_fold_destroy": (_I, [_I32]),
    "wfst_lattice_viterbi_workspace_bytes": (_Z, [_I, _I, _I]),
    "wfst_lattice_viterbi": (_I, [_P, _I, _I, _I, ctypes.POINTER(AcceptorBatch), _I, _P, _P, _P, _P, _Z, _P]),
    "wfst_graph_pack_sizes": (_I, [_P, _I, _P, _P, _P, _P, _P]),
    "wfst_graph_pack": (_I, [_P, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
}

_lock = threading.Lock()
_lib = None
_pytargets = None
PYTARGETS_PATH = os.path.join(_HERE, "lib", "lib_wfst_pytargets.so")


def pytargets():
    """The Python-list walker (csrc/pytargets.c), loaded with PyDLL (GIL held, borrowed
    references).  Built together with libwfst_b200.so; raises if it is missing."""
    global _pytargets
    if _pytargets is None:
        with _lock:
            if _pytargets is None:
                if not os.path.exists(PYTARGETS_PATH):
                    raise WfstError("lib_wfst_pytargets.so is not built (%s); run `make -C "
                                    "gtn_applications_b200/csrc`" % PYTARGETS_PATH)
                h = ctypes.PyDLL(PYTARGETS_PATH)
                h.wfst_pytargets_lengths.restype = ctypes.c_longlong
                h.wfst_pytargets_lengths.argtypes = [ctypes.py_object, ctypes.c_void_p, ctypes.c_longlong]
                h.wfst_pytargets_fill.restype = ctypes.c_longlong
                h.wfst_pytargets_fill.argtypes = [ctypes.py_object, ctypes.c_void_p, ctypes.c_longlong,
                                                  ctypes.c_void_p]
                _pytargets = h
    return _pytargets


class WfstError(RuntimeError):
    pass


def lib():
    """The loaded library; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise WfstError(
                    "libwfst_b200.so is not built (%s). Build it with "
                    "`python -c 'import __graft_entry__ as g; g.build()'` or "
                    "`make -C gtn_applications_b200/csrc`. There is no CPU fallback." % LIB_PATH)
            handle = ctypes.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(handle, name)  # AttributeError if the symbol is missing
                fn.restype = res
                fn.argtypes = args
            if handle.wfst_abi_version() != 2:
                raise WfstError("libwfst_b200.so ABI version mismatch")
            _lib = handle
    return _lib


def check(rc):
    if rc != 0:
        msg = lib().wfst_last_error().decode("utf-8", "replace")
        if rc == -1:
            raise ValueError(msg)
        raise WfstError("libwfst_b200 error %d: %s" % (rc, msg))


def launch_count():
    return int(lib().wfst_launch_count())
