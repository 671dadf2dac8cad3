def strtobool(val):
    v = str(val).lower()
    if v in ('y', 'yes', 't', 'true', 'on', '1'):
        return 1
    if v in ('n', 'no', 'f', 'false', 'off', '0'):
        return 0
    raise ValueError(f'invalid truth value {val!r}')
