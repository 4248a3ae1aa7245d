"""Importable alias of the ``ga-ddpg_b200/`` source directory (a hyphen is not a valid module name).

``import gaddpg_b200.agent`` resolves to ``ga-ddpg_b200/agent.py``: this package only extends its own
``__path__`` with that directory; all product code lives there.
"""
import os as _os

_SRC = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "ga-ddpg_b200")
__path__.insert(0, _SRC)
SRC_DIR = _SRC
