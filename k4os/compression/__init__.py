# K4os.Compression.* namespace
