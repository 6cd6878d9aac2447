from .triangulation import *
