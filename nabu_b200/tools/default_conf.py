"""Per-class default configuration merge (reference: nabu/tools/default_conf.py:9-36).

Every plugin class ships `<dir>/defaults/<classname.lower()>.cfg` with a [default] section; a key
missing from the recipe takes the default, and an *empty* default marks the key as mandatory.
A missing defaults file is tolerated (the reference's Recognizer has none)."""
import configparser
import os


def apply_defaults(conf, default_file):
    if not os.path.exists(default_file):
        return conf
    parser = configparser.ConfigParser()
    parser.read(default_file)
    for field, value in parser.items('default'):
        if field in conf:
            continue
        if value == '':
            raise Exception('the field %s was not found in the configuration file' % field)
        conf[field] = value
    return conf


def defaults_path(module_file, cls):
    return os.path.join(os.path.dirname(os.path.realpath(module_file)), 'defaults', cls.__name__.lower() + '.cfg')
