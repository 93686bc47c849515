"""Module named ``cars`` that shadows the reference's ``cars.py``: put this directory first on ``sys.path`` and every
reference driver (``animate.py``, ``environment.py``, ``learn.py``, ``artist.py``, ``test/profile_cars_update.py``)
gets the device-backed classes through its unchanged ``from cars import Cars, TrafficLights``."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.append(_ROOT)

from traffic_b200.facade import Cars, TrafficLights  # noqa: E402,F401
