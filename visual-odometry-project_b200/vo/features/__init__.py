from .klt import *
from .harris import *
from .sift import *
from .tracker import *
