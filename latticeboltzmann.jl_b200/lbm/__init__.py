"""lbm -- host-side mirror of the LatticeBoltzmann.jl API for the collide-stream-BC path,
running on liblbm_b200.so (hand-written sm_100a CUDA kernels behind a C ABI).

Import with `sys.path.insert(0, "<repo>/latticeboltzmann.jl_b200")` then `import lbm`.
Names follow src/LatticeBoltzmann.jl:31-76 (+ the unexported names the reference's tests,
benchmarks and notebooks use); Julia's `f!` is spelled `f_`.  Arrays are Float64 with shape
(NX, NY, Q) in Fortran order, i.e. exactly Julia's `f[x, y, i]` memory.
"""
from . import _abi
from ._abi import LbmError, pinned_empty
from .batch import BatchResult, simulate_many
from .boundary_conditions import (BoundaryCondition, BounceBack, Direction, East, MovingWall, North, South, West)
from .collision_models import (MRT, SRT, TRT, CollisionModel, IterativeInitializationCollisionModel, LatticeForce,
                               TRT_Lambda)
from .initial_conditions import (AnalyticalEquilibrium, AnalyticalEquilibriumAndOffEquilibrium, AnalyticalVelocity,
                                 AnalyticalVelocityAndStress, ConstantDensity, InitializationStrategy,
                                 IterativeInitialization, IterativeInitializationMeiEtAl,
                                 ZeroVelocityInitialCondition, initialize, initialize_mei_et_al, initialize_on_device)
from .model import (DeviceState, LatticeBoltzmannModel, apply_, apply_boundary_conditions_, clear_scratch_contexts,
                    collide_, collide_model_, next_model_, simulate, simulate_model, stream, stream_, stream_model_)
from .parallel import SlabComm, halo_rows_per_direction, slab_rows
from .problems import (CouetteFlow, DecayingShearFlow, FluidFlowProblem, LidDrivenCavityFlow,
                       LinearizedThermalDiffusion, LinearizedTransverseShearWave, PoiseuilleFlow, TGV,
                       TaylorGreenVortex, boundary_conditions, decay, decay_time, delta_t, delta_x, dimensionless_density,
                       dimensionless_force, dimensionless_pressure, dimensionless_stress, dimensionless_temperature,
                       dimensionless_velocity, dimensionless_viscosity, force, has_external_force, lattice_density,
                       lattice_force, lattice_pressure, lattice_temperature, lattice_velocity, lattice_viscosity, range_,
                       viscosity)
from .processing_methods import (CompareWithAnalyticalSolution, DensityConvergence, MeanVelocityStoppingCriteria,
                                 NoStoppingCriteria, ProcessIterativeInitialization, ProcessingMethod, StopCriteria, TakeSnapshots, TrackHydrodynamicErrors,
                                 VelocityConvergenceStoppingCriteria, process_)
from .quadratures import (D2Q4, D2Q5, D2Q9, D2Q13, D2Q17, D2Q21, D2Q37, Quadrature, Quadratures, dimension, opposite,
                          order)
from .vdf import (density, deviatoric_tensor, equilibrium, equilibrium_, equilibrium_coefficient, hermite,
                  hermite_based_equilibrium, momentum_flux, pressure, temperature, velocity, velocity_)

TRT_Λ = TRT_Lambda
