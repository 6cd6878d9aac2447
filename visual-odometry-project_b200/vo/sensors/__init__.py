from .camera import *
