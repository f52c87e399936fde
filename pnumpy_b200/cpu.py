"""Engine start-up and the enable/disable pair.  reference: ``src/pnumpy/cpu.py:90-149``.

(The reference file also carries a psutil-derived physical-core counter used only to pick a
worker count; lanes here are bounded by GPUs, so only the logical count is kept for reporting.)
"""
import os

from . import _pnumpy
from ._pnumpy import atop_enable, atop_disable, thread_enable, thread_disable

__all__ = ['cpu_count_linux', 'init', 'enable', 'disable']


def cpu_count_linux():
    """Return (logical CPUs, physical cores or None).  reference: ``src/pnumpy/cpu.py:60-88``."""
    try:
        logical = os.sysconf("SC_NPROCESSORS_ONLN")
    except (ValueError, OSError, AttributeError):
        logical = os.cpu_count()
    physical = None
    try:
        cores = set()
        base = "/sys/devices/system/cpu"
        for name in os.listdir(base):
            if name.startswith("cpu") and name[3:].isdigit():
                for leaf in ("core_cpus_list", "thread_siblings_list"):
                    path = os.path.join(base, name, "topology", leaf)
                    if os.path.exists(path):
                        with open(path) as f:
                            cores.add(f.read().strip())
                        break
        physical = len(cores) or None
    except OSError:
        pass
    return logical, physical


def init():
    """
    Called at load time to start the atop and dispatch engines: probes the GPUs, installs the
    hooks (``_pnumpy.initialize``).  Idempotent.  Raises ImportError when no B200 is usable.

    See Also
    --------
    pn.enable
    pn.disable
    """
    _pnumpy.initialize()


def enable():
    """
    Call to enable the atop engine and multi-lane dispatch (hooks stay installed either way).

    See Also
    --------
    pn.disable
    pn.atop_info
    """
    atop_enable()
    thread_enable()


def disable():
    """
    Call to disable the atop engine and multi-lane dispatch: every hooked loop runs NumPy's
    original inner loop.  The hooks are not removed (reference: SURVEY.md App. B).

    See Also
    --------
    pn.enable
    pn.atop_info
    """
    atop_disable()
    thread_disable()
