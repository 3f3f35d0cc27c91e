"""lambdapic_b200: the lambdaPIC per-step inner loop on B200, behind the reference's Simulation / Species /
callback-stage API.  ``import lambdapic_b200 as lambdapic`` is the intended drop-in for scripts
that stay on the accelerated path (periodic or CPML boundaries, laser antenna, moving window)."""
from .callback import Callback, callback
from .diagnostics import ExtractSpeciesDensity
from .laser import GaussianLaser2D, GaussianLaser3D, SimpleLaser2D, SimpleLaser3D
from .moving_window import MovingWindow
from .simulation import Simulation, Simulation2D, Simulation3D
from .species import Electron, Positron, Proton, Species
from .utils import SetTemperature, get_fields

c, e, epsilon_0, m_e, m_p, mu_0, pi = (299792458.0, 1.602176634e-19, 8.8541878188e-12, 9.1093837139e-31,
                                       1.67262192595e-27, 1.25663706127e-06, 3.141592653589793)

__all__ = ["Simulation", "Simulation2D", "Simulation3D", "Species", "Electron", "Proton", "Positron",
           "callback", "Callback", "GaussianLaser2D", "GaussianLaser3D", "SimpleLaser2D", "SimpleLaser3D", "MovingWindow", "ExtractSpeciesDensity", "SetTemperature", "get_fields", "c", "e", "epsilon_0", "m_e", "m_p", "mu_0", "pi"]
