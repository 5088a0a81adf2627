"""Drop-in for the reference's model_nerv.py (train_nerv_all.py:18 imports NeRV_Boost from here)."""
from bnerv_b200.models import NeRV_Boost  # noqa: F401
