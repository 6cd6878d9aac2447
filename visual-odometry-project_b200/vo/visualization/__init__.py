from .point_cloud import *
from .overlays import *
