"""placeholder (replaced below)"""
