"""see compat/matplotlib"""
