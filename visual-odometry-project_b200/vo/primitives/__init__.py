from .features import *
from .frame import *
from .matches import *
from .loader import *
from .state import *
