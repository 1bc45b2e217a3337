import zipfile, re, sys
import xml.etree.ElementTree as ET
NS={'m':'http://schemas.openxmlformats.org/spreadsheetml/2006/main'}
def read_xlsx(path):
    z=zipfile.ZipFile(path)
    ss=[]
    if 'xl/sharedStrings.xml' in z.namelist():
        r=ET.fromstring(z.read('xl/sharedStrings.xml'))
        for si in r.findall('m:si',NS):
            ss.append(''.join(t.text or '' for t in si.iter('{%s}t'%NS['m'])))
    wb=ET.fromstring(z.read('xl/workbook.xml'))
    rels=ET.fromstring(z.read('xl/_rels/workbook.xml.rels'))
    rid={r.get('Id'):r.get('Target') for r in rels}
    out={}
    for s in wb.find('m:sheets',NS):
        name=s.get('name'); r=s.get('{http://schemas.openxmlformats.org/officeDocument/2006/relationships}id')
        tgt=rid[r]; tgt='xl/'+tgt if not tgt.startswith('/') else tgt[1:]
        sh=ET.fromstring(z.read(tgt))
        rows=[]
        for row in sh.iter('{%s}row'%NS['m']):
            cells={}
            for c in row.findall('m:c',NS):
                ref=c.get('r'); col=re.match(r'[A-Z]+',ref).group(0)
                ci=0
                for ch in col: ci=ci*26+ord(ch)-64
                v=c.find('m:v',NS); t=c.get('t')
                if v is None:
                    is_=c.find('m:is',NS)
                    val=''.join(x.text or '' for x in is_.iter('{%s}t'%NS['m'])) if is_ is not None else None
                elif t=='s': val=ss[int(v.text)]
                else: val=v.text
                cells[ci-1]=val
            if cells:
                n=max(cells)+1
                rows.append([cells.get(i) for i in range(n)])
            else: rows.append([])
        out[name]=rows
    return out
if __name__=='__main__':
    d=read_xlsx(sys.argv[1])
    for k,v in d.items():
        print('==',k,len(v))
        for r in v[:int(sys.argv[2]) if len(sys.argv)>2 else 10]: print(r)
