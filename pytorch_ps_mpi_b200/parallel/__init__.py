"""Parallel engines: host transports, flat arena layout, symmetric memory, device engine."""
