from .ransac import *
