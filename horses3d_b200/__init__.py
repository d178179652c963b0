"""horses3d_b200: B200-native explicit compressible-NS residual + RK step behind the HORSES3D interfaces."""
