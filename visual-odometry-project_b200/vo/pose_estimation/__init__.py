from .p3p import *
