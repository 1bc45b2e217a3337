"""heros-b200: B200-native engine for the transmission + disease-progression hot path of MedLabs/HEROS.

Layout: ``csrc/`` hand-written sm_100a CUDA kernels + the C ABI (``include/heros_gpu.h``); ``_lib`` ctypes binding;
``engine`` numpy wrapper; ``disease`` / ``policy`` mirror of the reference's plug-in classes; ``properties``
parameter parsing; ``scenario`` synthetic populations.  Importing the package does not need a GPU; creating an
engine does, and there is no CPU fallback.
"""
from . import _lib, properties  # noqa: F401
from ._lib import (HerosError, MODEL_AREA, MODEL_DISTANCE, PHASE_NAMES, TRIGGER_ENTER_AND_EXIT,  # noqa: F401
                   TRIGGER_EXIT_ONLY)
from .build import build  # noqa: F401
from .disease import (Covid19Progression, Covid19TransmissionArea, Covid19TransmissionDistance,  # noqa: F401
                      DiseaseTransmission, HerosModel, InfectionRecord, construct_disease_model)
from .engine import HerosEngine  # noqa: F401
from .monitor import DiseaseMonitor, LocationDump, compartment_series  # noqa: F401
from .policy import (DiseasePolicy, LocationPolicy, apply_location_policies, schedule_disease_policies,  # noqa: F401
                     schedule_location_policies, schedule_location_policy_rows)
from . import sharded  # noqa: F401

__version__ = "0.1.0"
