"""Drop-in for the reference's model_hnerv.py (train_nerv_all.py:16 imports HNeRV, HNeRVDecoder, HNeRV_Boost)."""
from bnerv_b200.models import HNeRV, HNeRVDecoder, HNeRV_Boost  # noqa: F401
