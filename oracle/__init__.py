"""CPU restatement ("oracle") of the Faster R-CNN hot path of andreaskoepf/faster-rcnn.torch.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import
or execute it, and only as the checker / the timed CPU baseline.

PARITY UNPINNED: the reference ships no tests, no golden vectors and cannot be executed in this
environment (no Lua/LuaJIT/Torch7; its CUDA path needs 2015 cunn/cutorch).  The restatement follows the
reference sources line by line (every function cites file:line) and is pinned only by
  * the hand-derived known-answer values of SURVEY.md section 8(c) (an independent trace), and
  * PyTorch-CPU fp32 ops as the direct descendants of TH/THNN for conv / pool / linear numerics.
"""
from .rect import Rect  # noqa: F401
from .localizer import Localizer  # noqa: F401
from .anchors import Anchors  # noqa: F401
