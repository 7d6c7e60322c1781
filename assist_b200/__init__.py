"""assist-b200: batched test-particle integration (ASSIST forces + IAS15) on NVIDIA B200.

`Ephem`, `Extras`, `ASSIST_BODY_IDS`, `ASSIST_FORCES` mirror the reference's `assist` Python package;
`Simulation` / `Particle` stand in for the `rebound` package on that path; `Batch` is the many-particle API
(assist_gpu.h).  The shared library is loaded on first use and there is no CPU fallback.
"""
from .api import (ASSIST_BODY_IDS, ASSIST_FORCES, Ephem, Extras, Particle, Simulation, SimulationArchive,  # noqa: F401
                  assist_create_interpolated_simulation, assist_error_messages, simulation_convert_to_rebound)
from .batch import Batch, EphemHandle  # noqa: F401

__version__ = "0.1.0"
