import json,sys
a=json.load(open(sys.argv[1])); b=json.load(open(sys.argv[2]))
for x,y in zip(a['full']['step_log'], b['full']['step_log']):
    print(x[1], x[3:], y[3:], abs(x[2]-y[2])/abs(y[2]))
