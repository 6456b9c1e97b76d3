"""sonar_3d_reconstruction_b200 -- B200-native (sm_100a) implementation of the per-frame
sonar -> voxel log-odds hot path of luckkim123/sonar_3d_reconstruction, behind the reference's
own SonarTo3DMapper / SimpleOctree Python API.  See DESIGN.md and INTEGRATION.md."""
from .mapper import SimpleOctree, SonarTo3DMapper

__all__ = ["SimpleOctree", "SonarTo3DMapper"]
__version__ = "0.1.0"
