# namespace root mirroring the reference's K4os.* assemblies
