# Data module
